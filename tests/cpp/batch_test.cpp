// C++ port of the reference's tests/batch_test.rs against include/bdeflate.hpp
// (the host mirror of src/batch.rs).  Built and run by tests/test_gpu_cpp_host.py.
#include <cstdio>
#include <cstdlib>
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
    std::puts("batch_test.cpp: all reference batch tests passed");
    return 0;
}
