// bdeflate.hpp — C++ host mirror of the reference's batch API over the C ABI.
//
// Same names, argument meaning and failure behaviour as reference src/batch.rs:
//   BatchCompressor::new(level) / compress_batch(&[&[u8]]) -> Vec<Vec<u8>>        (:12,:20-58)
//       a stream that fails (output above its bound) yields an EMPTY vector        (:52-53)
//   BatchDecompressor::new() / decompress_batch(&[&[u8]], &[usize]) -> Vec<Option<Vec<u8>>> (:70,:74-101)
//       a stream that fails yields std::nullopt                                     (:95-96)
//       result length = min(inputs, max_out_sizes)  (zip, :79-81)
// plus the `format` selector of the C ABI, the safe API of src/api.rs (Compressor / Decompressor)
// and the stream adapters of src/stream.rs (DeflateEncoder, DeflateDecoder).  Header-only; link with libbdeflate.so.
// A call-level failure (no GPU, CUDA error, unsupported request) throws — there
// is no CPU fallback behind this API.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "bdeflate.h"

namespace bdf {

using Bytes = std::vector<uint8_t>;
using ByteView = std::pair<const uint8_t *, size_t>;

class Context {
public:
    explicit Context(int device = 0)
    {
        if (bdf_ctx_create(device, &ctx_) != BDF_E_OK)
            throw std::runtime_error("bdf_ctx_create failed: no usable CUDA device (no CPU fallback)");
    }
    ~Context() { bdf_ctx_destroy(ctx_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    bdf_ctx *get() const { return ctx_; }
    void check(int rc) const
    {
        if (rc != BDF_E_OK) throw std::runtime_error(std::string("libbdeflate: ") + bdf_last_error(ctx_));
    }

private:
    bdf_ctx *ctx_ = nullptr;
};

namespace detail {
inline void flatten(const std::vector<ByteView> &in, size_t n, Bytes &flat, std::vector<uint64_t> &off)
{
    off.assign(n + 1, 0);
    for (size_t i = 0; i < n; i++) off[i + 1] = off[i] + in[i].second;
    flat.resize(off[n] ? off[n] : 1);
    for (size_t i = 0; i < n; i++)
        if (in[i].second) std::memcpy(flat.data() + off[i], in[i].first, in[i].second);
}
}  // namespace detail

class BatchCompressor {
public:
    explicit BatchCompressor(size_t level, int format = BDF_RAW, std::shared_ptr<Context> ctx = nullptr)
        : level_((int)(level > 12 ? 12 : level)), format_(format), ctx_(ctx ? ctx : std::make_shared<Context>())
    {
    }
    std::vector<Bytes> compress_batch(const std::vector<ByteView> &inputs) const
    {
        const size_t n = inputs.size();
        std::vector<Bytes> res(n);
        if (n == 0) return res;
        Bytes flat;
        std::vector<uint64_t> in_off, out_off(n), out_size(n);
        std::vector<int32_t> status(n);
        detail::flatten(inputs, n, flat, in_off);
        uint64_t total = 0;
        for (size_t i = 0; i < n; i++) {
            out_off[i] = total;
            total += bdf_compress_bound(format_, inputs[i].second);   // src/batch.rs:39
        }
        Bytes slab(total ? total : 1);
        ctx_->check(bdf_compress_batch_host(ctx_->get(), level_, format_, flat.data(), in_off.data(), n,
                                            slab.data(), out_off.data(), out_size.data(), status.data()));
        for (size_t i = 0; i < n; i++)
            if (status[i] == BDF_OK)
                res[i].assign(slab.begin() + out_off[i], slab.begin() + out_off[i] + out_size[i]);
        return res;
    }
    // Compressor::compress_to_size(input, final_block) per buffer (src/compress/mod.rs:1073-1094):
    // estimated raw-DEFLATE bytes, no output slab; buffers of at most 256 KiB
    std::vector<size_t> compress_to_size_batch(const std::vector<ByteView> &inputs, bool final_block = true) const
    {
        const size_t n = inputs.size();
        std::vector<size_t> res(n);
        if (n == 0) return res;
        Bytes flat;
        std::vector<uint64_t> in_off, out_size(n);
        std::vector<int32_t> status(n);
        detail::flatten(inputs, n, flat, in_off);
        ctx_->check(bdf_compress_size_batch_host(ctx_->get(), level_, flat.data(), in_off.data(), n,
                                                 final_block ? 1 : 0, out_size.data(), status.data()));
        for (size_t i = 0; i < n; i++) {
            if (status[i] != BDF_OK) throw std::runtime_error("compress_to_size: buffer not supported");
            res[i] = (size_t)out_size[i];
        }
        return res;
    }

private:
    int level_, format_;
    std::shared_ptr<Context> ctx_;
};

class BatchDecompressor {
public:
    explicit BatchDecompressor(int format = BDF_RAW, std::shared_ptr<Context> ctx = nullptr)
        : format_(format), ctx_(ctx ? ctx : std::make_shared<Context>())
    {
    }
    std::vector<std::optional<Bytes>> decompress_batch(const std::vector<ByteView> &inputs,
                                                       const std::vector<size_t> &max_out_sizes) const
    {
        const size_t n = inputs.size() < max_out_sizes.size() ? inputs.size() : max_out_sizes.size();
        std::vector<std::optional<Bytes>> res(n);
        if (n == 0) return res;
        Bytes flat;
        std::vector<uint64_t> in_off, out_off(n), max_out(n), out_size(n);
        std::vector<int32_t> status(n);
        detail::flatten(inputs, n, flat, in_off);
        uint64_t total = 0;
        for (size_t i = 0; i < n; i++) {
            out_off[i] = total;
            max_out[i] = max_out_sizes[i];
            total += max_out[i];
        }
        Bytes slab(total ? total : 1);
        ctx_->check(bdf_decompress_batch_host(ctx_->get(), format_, flat.data(), in_off.data(), n, slab.data(),
                                              out_off.data(), max_out.data(), out_size.data(), nullptr,
                                              status.data()));
        for (size_t i = 0; i < n; i++)
            if (status[i] == BDF_OK)
                res[i] = Bytes(slab.begin() + out_off[i], slab.begin() + out_off[i] + out_size[i]);
        return res;
    }

private:
    int format_;
    std::shared_ptr<Context> ctx_;
};

// ---------------------------------------------------------------------------------------------
// Safe API of the reference (src/api.rs): same names, argument meaning and error behaviour.
//   Compressor::new(level): level outside 0..=12 -> InvalidInput                      (:10-16)
//   Decompressor: expected_size > len * limit_ratio + 4096 -> InvalidInput "safety limit",
//                 expected_size > max_memory_limit -> InvalidInput "maximum memory limit" (:213-239)
//   failures of the codec itself -> InvalidData / "Compression failed"
// InvalidInput maps to std::invalid_argument, the others to std::runtime_error.
class Compressor {
public:
    explicit Compressor(int level, std::shared_ptr<Context> ctx = nullptr) : level_(level), ctx_(std::move(ctx))
    {
        if (level < 0 || level > 12) throw std::invalid_argument("Compression level must be between 0 and 12");
    }
    Bytes compress_deflate(ByteView data) { return one(BDF_RAW, data); }
    Bytes compress_zlib(ByteView data) { return one(BDF_ZLIB, data); }
    Bytes compress_gzip(ByteView data) { return one(BDF_GZIP, data); }
    size_t deflate_compress_bound(size_t n) const { return bdf_compress_bound(BDF_RAW, n); }
    size_t zlib_compress_bound(size_t n) const { return bdf_compress_bound(BDF_ZLIB, n); }
    size_t gzip_compress_bound(size_t n) const { return bdf_compress_bound(BDF_GZIP, n); }

private:
    Bytes one(int format, ByteView data)
    {
        if (!ctx_) ctx_ = std::make_shared<Context>();
        auto out = BatchCompressor((size_t)level_, format, ctx_).compress_batch({data});
        if (out[0].empty() && !(level_ == 0 && data.second == 0 && format == BDF_RAW))
            throw std::runtime_error("Compression failed");
        return std::move(out[0]);
    }
    int level_;
    std::shared_ptr<Context> ctx_;
};

class Decompressor {
public:
    explicit Decompressor(std::shared_ptr<Context> ctx = nullptr) : ctx_(std::move(ctx)) {}
    void set_max_memory_limit(size_t limit) { max_memory_limit_ = limit; }
    void set_limit_ratio(size_t ratio) { limit_ratio_ = ratio; }
    Bytes decompress_deflate(ByteView data, size_t expected_size) { return one(BDF_RAW, data, expected_size); }
    Bytes decompress_zlib(ByteView data, size_t expected_size) { return one(BDF_ZLIB, data, expected_size); }
    Bytes decompress_gzip(ByteView data, size_t expected_size) { return one(BDF_GZIP, data, expected_size); }

private:
    static size_t sat_mul(size_t a, size_t b) { return b && a > SIZE_MAX / b ? SIZE_MAX : a * b; }
    static size_t sat_add(size_t a, size_t b) { return a > SIZE_MAX - b ? SIZE_MAX : a + b; }
    Bytes one(int format, ByteView data, size_t expected_size)
    {
        const size_t limit = sat_add(sat_mul(data.second, limit_ratio_), 4096);
        if (expected_size > limit)
            throw std::invalid_argument("Expected size " + std::to_string(expected_size) +
                                        " exceeds safety limit for input size " + std::to_string(data.second));
        if (expected_size > max_memory_limit_)
            throw std::invalid_argument("Expected size " + std::to_string(expected_size) +
                                        " exceeds maximum memory limit " + std::to_string(max_memory_limit_));
        if (!ctx_) ctx_ = std::make_shared<Context>();
        auto out = BatchDecompressor(format, ctx_).decompress_batch({data}, {expected_size});
        if (!out[0]) throw std::runtime_error("Decompression failed");
        return std::move(*out[0]);
    }
    size_t max_memory_limit_ = SIZE_MAX, limit_ratio_ = 2000;
    std::shared_ptr<Context> ctx_;
};

// ---------------------------------------------------------------------------------------------
// DeflateEncoder of the reference (src/stream.rs:15-240): bytes are buffered (1 MiB by default), a
// full buffer is cut into 256 KiB chunks, every chunk goes through a fresh compressor and ends in
// a sync flush, the last chunk of finish() ends the stream.  All chunks of a buffer are ONE
// bdf_compress_units_host call.  `Sink` needs write(const uint8_t*, size_t) and flush().
template <class Sink>
class DeflateEncoder {
public:
    DeflateEncoder(Sink &sink, size_t level, std::shared_ptr<Context> ctx = nullptr)
        : sink_(&sink), level_((int)(level > 12 ? 12 : level)), ctx_(ctx ? ctx : std::make_shared<Context>())
    {
        buffer_.reserve(buffer_size_);
    }
    ~DeflateEncoder()
    {
        if (sink_) try { flush_buffer(true); } catch (...) {}        // Drop, :234-240: errors are ignored
    }
    DeflateEncoder &with_buffer_size(size_t n) { buffer_size_ = n; return *this; }
    size_t write(const uint8_t *p, size_t n)
    {
        buffer_.insert(buffer_.end(), p, p + n);
        if (buffer_.size() >= buffer_size_) flush_buffer(false);
        return n;
    }
    void flush() { flush_buffer(false); sink_->flush(); }
    void finish() { flush_buffer(true); sink_ = nullptr; }

private:
    void flush_buffer(bool final_block)
    {
        if (buffer_.empty() && !final_block) return;
        constexpr size_t CHUNK = 256 * 1024;
        const size_t n = buffer_.empty() ? 1 : (buffer_.size() + CHUNK - 1) / CHUNK;
        std::vector<uint64_t> off(n + 1), out_off(n), out_size(n);
        std::vector<uint8_t> mode(n, (uint8_t)BDF_FLUSH_SYNC);
        std::vector<int32_t> status(n);
        uint64_t total = 0;
        for (size_t i = 0; i < n; i++) {
            off[i] = i * CHUNK;
            const size_t len = std::min(buffer_.size() - off[i], CHUNK);
            out_off[i] = total;
            total += bdf_compress_bound(BDF_RAW, len) + 5;
        }
        off[n] = buffer_.size();
        if (final_block) mode[n - 1] = (uint8_t)BDF_FLUSH_FINISH;
        Bytes slab(total), dummy(1);
        ctx_->check(bdf_compress_units_host(ctx_->get(), level_, buffer_.empty() ? dummy.data() : buffer_.data(),
                                            off.data(), mode.data(), n, slab.data(), out_off.data(),
                                            out_size.data(), status.data()));
        for (size_t i = 0; i < n; i++) {
            if (status[i] != BDF_OK) throw std::runtime_error("Compression failed");
            sink_->write(slab.data() + out_off[i], out_size[i]);
        }
        buffer_.clear();
    }
    Sink *sink_;
    int level_;
    size_t buffer_size_ = 1024 * 1024;
    Bytes buffer_;
    std::shared_ptr<Context> ctx_;
};

// ---------------------------------------------------------------------------------------------
// DeflateDecoder of the reference (src/stream.rs:243-376): an incremental reader over a raw DEFLATE
// stream.  A 64 KiB window whose lower half is match history, input pulled from `Source` only when
// the decoder asks for it, the decoder resumed where the last read left it through
// bdf_inflate_resume_batch_host (one decoder state here; the call advances many per launch).
// `Source` needs size_t read(uint8_t *, size_t) returning 0 at the end of input.
// Errors: std::runtime_error("deflate decompression failed") (InvalidData, :330-340) and
// std::runtime_error("unexpected EOF") (:353-366).
template <class Source>
class DeflateDecoder {
public:
    explicit DeflateDecoder(Source &src, std::shared_ptr<Context> ctx = nullptr)
        : src_(&src), ctx_(ctx ? ctx : std::make_shared<Context>()), window_(WINDOW)
    {
        std::memset(&state_, 0, sizeof(state_));
    }
    // up to n bytes, 0 only at the end of the stream
    size_t read(uint8_t *buf, size_t n)
    {
        if (n == 0) return 0;
        fill();
        const size_t count = std::min<size_t>(n, write_pos_ - read_pos_);
        std::memcpy(buf, window_.data() + read_pos_, count);
        read_pos_ += count;
        return count;
    }
    Bytes read_to_end()
    {
        Bytes out, piece(WINDOW);
        for (;;) {
            const size_t k = read(piece.data(), piece.size());
            if (k == 0) return out;
            out.insert(out.end(), piece.begin(), piece.begin() + k);
        }
    }

private:
    static constexpr size_t HISTORY = 32 * 1024, WINDOW = 64 * 1024 + 258, INPUT_CHUNK = 32 * 1024;
    void fill()
    {
        while (read_pos_ == write_pos_ && !done_) {
            if (WINDOW - write_pos_ < 258 && write_pos_ > HISTORY) {     // everything was read: keep the history (:284-295)
                std::memmove(window_.data(), window_.data() + (write_pos_ - HISTORY), HISTORY);
                write_pos_ = read_pos_ = HISTORY;
            }
            const uint8_t dummy = 0;
            const uint64_t in_off[2] = {0, input_.size()}, win_off = 0, win_cap = WINDOW;
            const uint8_t fin = in_final_ ? 1 : 0;
            uint64_t pos = write_pos_, used = 0;
            int32_t status = 0;
            ctx_->check(bdf_inflate_resume_batch_host(ctx_->get(), 1, &state_, input_.empty() ? &dummy : input_.data(), in_off,
                                                      &fin, window_.data(), &win_off, &win_cap, &pos, &used, &status));
            input_.erase(input_.begin(), input_.begin() + (size_t)used);
            write_pos_ = (size_t)pos;
            if (status == BDF_OK) { done_ = true; break; }
            if (status == BDF_BAD_DATA) throw std::runtime_error("deflate decompression failed");
            if (write_pos_ > read_pos_ || status == BDF_INSUFFICIENT_SPACE) continue;
            if (in_final_) throw std::runtime_error("unexpected EOF");
            const size_t have = input_.size();
            input_.resize(have + INPUT_CHUNK);
            const size_t got = src_->read(input_.data() + have, INPUT_CHUNK);
            input_.resize(have + got);
            if (got == 0) {
                if (input_.empty() && state_.phase == 0) done_ = true;      // the input ends between two blocks: Ok(0), :367-369
                else in_final_ = true;
            }
        }
    }
    Source *src_;
    std::shared_ptr<Context> ctx_;
    bdf_inflate_state state_;
    Bytes window_, input_;
    size_t read_pos_ = 0, write_pos_ = 0;
    bool in_final_ = false, done_ = false;
};

}  // namespace bdf
