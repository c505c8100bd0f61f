// C++ port of the reference's tests/batch_test.rs against include/bdeflate.hpp
// (the host mirror of src/batch.rs).  Built and run by tests/test_gpu_cpp_host.py.
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "bdeflate.hpp"

#define CHECK(cond)                                                      \
    do {                                                                 \
        if (!(cond)) {                                                   \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            std::exit(1);                                                \
        }                                                                \
    } while (0)

static bdf::ByteView view(const std::string &s) { return {reinterpret_cast<const uint8_t *>(s.data()), s.size()}; }

int main()
{
    auto ctx = std::make_shared<bdf::Context>(0);
    // test_batch_compress_decompress_roundtrip (tests/batch_test.rs:4-39)
    {
        std::vector<std::string> data = {
            "Hello world! This is a test string for deflate compression.", "Another test string.",
            "Repeating pattern repeating pattern repeating pattern repeating pattern.", "Short",
            std::string(1000, '\0')};
        std::vector<bdf::ByteView> in;
        std::vector<size_t> sizes;
        for (auto &s : data) { in.push_back(view(s)); sizes.push_back(s.size()); }
        bdf::BatchCompressor c(6, BDF_RAW, ctx);
        auto comp = c.compress_batch(in);
        CHECK(comp.size() == in.size());
        std::vector<bdf::ByteView> cin;
        for (auto &v : comp) cin.push_back({v.data(), v.size()});
        bdf::BatchDecompressor d(BDF_RAW, ctx);
        auto out = d.decompress_batch(cin, sizes);
        CHECK(out.size() == in.size());
        for (size_t i = 0; i < in.size(); i++) {
            CHECK(out[i].has_value());
            CHECK(std::string(out[i]->begin(), out[i]->end()) == data[i]);
        }
        // test_batch_insufficient_space (:86-100): max_out = len - 1 -> None
        std::vector<size_t> small = sizes;
        for (auto &s : small) s -= 1;
        auto out2 = d.decompress_batch(cin, small);
        for (auto &o : out2) CHECK(!o.has_value());
    }
    // test_batch_empty (:42-50)
    {
        bdf::BatchCompressor c(6, BDF_RAW, ctx);
        CHECK(c.compress_batch({}).empty());
        bdf::BatchDecompressor d(BDF_RAW, ctx);
        CHECK(d.decompress_batch({}, {}).empty());
    }
    // test_batch_empty_input (:53-70)
    {
        std::string a, b = "Not empty";
        bdf::BatchCompressor c(6, BDF_RAW, ctx);
        auto comp = c.compress_batch({view(a), view(b)});
        CHECK(comp.size() == 2 && !comp[0].empty());
        bdf::BatchDecompressor d(BDF_RAW, ctx);
        auto out = d.decompress_batch({{comp[0].data(), comp[0].size()}, {comp[1].data(), comp[1].size()}}, {0, 9});
        CHECK(out[0].has_value() && out[0]->empty());
        CHECK(out[1].has_value() && std::string(out[1]->begin(), out[1]->end()) == b);
    }
    // test_batch_decompress_bad_data (:73-83)
    {
        const uint8_t bad[6] = {0, 1, 2, 3, 4, 5};
        bdf::BatchDecompressor d(BDF_RAW, ctx);
        auto out = d.decompress_batch({{bad, 6}}, {100});
        CHECK(out.size() == 1 && !out[0].has_value());
    }
    // safe API: tests/security_limit.rs:5-57,60-107 and the level check of api.rs:10-16
    {
        bool threw = false;
        try { bdf::Compressor bad(13); } catch (const std::invalid_argument &) { threw = true; }
        CHECK(threw);
        bdf::Decompressor d(ctx);
        const std::vector<uint8_t> zeros(10, 0);
        threw = false;
        try { d.decompress_deflate({zeros.data(), zeros.size()}, 1000000); }
        catch (const std::invalid_argument &e) { threw = std::string(e.what()).find("safety limit") != std::string::npos; }
        CHECK(threw);
        d.set_max_memory_limit(50u << 20);
        const std::vector<uint8_t> mb(1 << 20, 0);
        threw = false;
        try { d.decompress_deflate({mb.data(), mb.size()}, 100u << 20); }
        catch (const std::invalid_argument &e) { threw = std::string(e.what()).find("maximum memory limit") != std::string::npos; }
        CHECK(threw);
        threw = false;                                   // within the limits: garbage is InvalidData, not InvalidInput
        try { bdf::Decompressor d2(ctx); d2.decompress_deflate({zeros.data(), zeros.size()}, 20000); }
        catch (const std::invalid_argument &) { CHECK(false); }
        catch (const std::runtime_error &) { threw = true; }
        CHECK(threw);
        std::string original;
        for (int i = 0; i < 10; i++) original += "Hello world";
        bdf::Compressor c(1, ctx);
        auto comp = c.compress_deflate(view(original));
        bdf::Decompressor d3(ctx);
        d3.set_max_memory_limit(1 << 20);
        auto back = d3.decompress_deflate({comp.data(), comp.size()}, original.size());
        CHECK(std::string(back.begin(), back.end()) == original);
        auto z = c.compress_zlib(view(original));
        auto zb = d3.decompress_zlib({z.data(), z.size()}, original.size());
        CHECK(std::string(zb.begin(), zb.end()) == original);
    }
    // stream encoder: tests/stream_test.rs:41-54 (round trip), tests/buffer_size_test.rs:24-58
    {
        struct VecSink {
            std::vector<uint8_t> data;
            int flushes = 0;
            void write(const uint8_t *p, size_t n) { data.insert(data.end(), p, p + n); }
            void flush() { flushes++; }
        };
        std::vector<uint8_t> data(700000);
        for (size_t i = 0; i < data.size(); i++) data[i] = (uint8_t)((i * 7 + i / 1000) % 251);
        VecSink sink;
        {
            bdf::DeflateEncoder<VecSink> enc(sink, 6, ctx);
            enc.with_buffer_size(300000);
            enc.write(data.data(), 150000);
            CHECK(sink.data.empty());
            enc.write(data.data() + 150000, 250000);          // 400000 >= 300000: two chunks go out, sync-flushed
            const size_t n1 = sink.data.size();
            CHECK(n1 > 0);
            enc.write(data.data() + 400000, 300000);
            enc.flush();
            CHECK(sink.flushes == 1);
            enc.finish();
            CHECK(sink.data.size() > n1);
        }
        bdf::BatchDecompressor d(BDF_RAW, ctx);
        auto out = d.decompress_batch({{sink.data.data(), sink.data.size()}}, {data.size()});
        CHECK(out[0].has_value() && *out[0] == data);
        // stream decoder: tests/stream_test.rs:41-54 reads the encoder's output back in small pieces
        struct VecSource {
            const std::vector<uint8_t> *v;
            size_t pos = 0, piece;
            size_t read(uint8_t *p, size_t n)
            {
                const size_t k = std::min({n, piece, v->size() - pos});
                std::memcpy(p, v->data() + pos, k);
                pos += k;
                return k;
            }
        };
        VecSource src{&sink.data, 0, 5000};
        bdf::DeflateDecoder<VecSource> dec(src, ctx);
        std::vector<uint8_t> back, piece(1000);
        for (;;) {
            const size_t k = dec.read(piece.data(), piece.size());
            if (k == 0) break;
            back.insert(back.end(), piece.begin(), piece.begin() + k);
        }
        CHECK(back == data);
        // a truncated stream is an error, not a short result
        std::vector<uint8_t> cut(sink.data.begin(), sink.data.begin() + sink.data.size() / 2);
        VecSource src2{&cut, 0, 70000};
        bdf::DeflateDecoder<VecSource> dec2(src2, ctx);
        bool threw = false;
        try { dec2.read_to_end(); } catch (const std::runtime_error &) { threw = true; }
        CHECK(threw);
    }
    // size estimation (Compressor::compress_to_size, src/compress/mod.rs:1073-1094): at the greedy /
    // lazy levels the estimate is the length of the stream the compressor writes
    {
        std::vector<uint8_t> data(50000);
        for (size_t i = 0; i < data.size(); i++) data[i] = (uint8_t)((i * 13 + i / 700) % 241);
        bdf::BatchCompressor c6(6, BDF_RAW, ctx);
        auto est = c6.compress_to_size_batch({{data.data(), data.size()}, {data.data(), 0}});
        auto comp = c6.compress_batch({{data.data(), data.size()}});
        CHECK(est[0] == comp[0].size());
        CHECK(est[1] == 0);
        bdf::BatchCompressor c0(0, BDF_RAW, ctx);
        CHECK(c0.compress_to_size_batch({{data.data(), 0}}, true)[0] == 5);
        CHECK(c0.compress_to_size_batch({{data.data(), 0}}, false)[0] == 0);
        CHECK(c0.compress_to_size_batch({{data.data(), data.size()}})[0] == data.size() + 5);
    }
    std::puts("batch_test.cpp: all reference batch tests passed");
    return 0;
}
