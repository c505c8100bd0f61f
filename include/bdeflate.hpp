// bdeflate.hpp — C++ host mirror of the reference's batch API over the C ABI.
//
// Same names, argument meaning and failure behaviour as reference src/batch.rs:
//   BatchCompressor::new(level) / compress_batch(&[&[u8]]) -> Vec<Vec<u8>>        (:12,:20-58)
//       a stream that fails (output above its bound) yields an EMPTY vector        (:52-53)
//   BatchDecompressor::new() / decompress_batch(&[&[u8]], &[usize]) -> Vec<Option<Vec<u8>>> (:70,:74-101)
//       a stream that fails yields std::nullopt                                     (:95-96)
//       result length = min(inputs, max_out_sizes)  (zip, :79-81)
// plus the `format` selector of the C ABI.  Header-only; link with libbdeflate.so.
// A call-level failure (no GPU, CUDA error, unsupported request) throws — there
// is no CPU fallback behind this API.
#pragma once
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

}  // namespace bdf
