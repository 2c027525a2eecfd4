// Mean-squared-error loss and its gradient in ONE streaming kernel (the metric is
// fwd+bwd of loss = mean((img - target)^2), the loss of the reference's
// examples/single_image.py:21-30 and examples/benchmark.py).  An eager host framework spends
// three kernels and 150 MB of traffic on it at 1920x1080 (subtract, dot, scale); this reads the
// image and the target once and writes the cotangent once (75 MB):
//   v_img[i] = (img[i] - target[i]) * scale          (scale = 2 / n from the caller, one rounding
//                                                      each: the eager result bit for bit)
//   loss     = sum_i (img[i] - target[i])^2 / n       (per-CTA partial sums in a per-stream
//                                                      scratch; the CTA that draws the last ticket
//                                                      adds them in CTA order -- the same bits on
//                                                      every run, no memset in front of the kernel;
//                                                      atomicInc wraps the ticket back to 0)
#include "../../include/jaxsplat_b200.h"
#include "common.cuh"

#include <map>
#include <mutex>
#include <utility>

namespace jsb {
namespace {

__device__ __forceinline__ float sq_diff(float a, float b, float scale, float &v) {
    const float d = __fsub_rn(a, b);
    v = __fmul_rn(d, scale);
    return d * d;
}

// target element i as float: float targets as they are, 8-bit ones as float(t) * (1 / 255) -- the
// decode an eager pipeline runs as its own kernel (u8 -> f32 copy, scale) before the loss
__device__ __forceinline__ float target_at(const float *t, size_t i) { return t[i]; }
__device__ __forceinline__ float target_at(const unsigned char *t, size_t i) {
    return __fmul_rn((float)t[i], 1.0f / 255.0f);
}
__device__ __forceinline__ float4 target4_at(const float *t, size_t i) {
    return __ldg(reinterpret_cast<const float4 *>(t) + i);
}
__device__ __forceinline__ float4 target4_at(const unsigned char *t, size_t i) {
    const uchar4 q = __ldg(reinterpret_cast<const uchar4 *>(t) + i);
    return make_float4(__fmul_rn((float)q.x, 1.0f / 255.0f), __fmul_rn((float)q.y, 1.0f / 255.0f),
                       __fmul_rn((float)q.z, 1.0f / 255.0f), __fmul_rn((float)q.w, 1.0f / 255.0f));
}

template <typename TargetT>
__global__ void __launch_bounds__(256)
mse_loss_grad_kernel(const float *__restrict__ img, const TargetT *__restrict__ target, size_t n,
                     float scale, float inv_n, float *__restrict__ v_img, float *__restrict__ loss,
                     float *__restrict__ partials, unsigned *__restrict__ ticket, bool vec) {
    __shared__ float s_part[256 / 32];
    __shared__ bool s_last;
    chain_wait();
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    float acc = 0.f;
    size_t done = 0;
    if (vec) {
        const size_t nv = n >> 2;
        for (size_t i = tid; i < nv; i += stride) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(img) + i);
            const float4 b = target4_at(target, i);
            float4 v;
            acc += sq_diff(a.x, b.x, scale, v.x);
            acc += sq_diff(a.y, b.y, scale, v.y);
            acc += sq_diff(a.z, b.z, scale, v.z);
            acc += sq_diff(a.w, b.w, scale, v.w);
            reinterpret_cast<float4 *>(v_img)[i] = v;
        }
        done = nv << 2;
    }
    for (size_t i = done + tid; i < n; i += stride) {
        float v;
        acc += sq_diff(img[i], target_at(target, i), scale, v);
        v_img[i] = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float total = 0.f;
#pragma unroll
        for (int w = 0; w < 256 / 32; ++w) total += s_part[w];
        __stcg(partials + blockIdx.x, total * inv_n);
        __threadfence();
        s_last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;  // (wraps to 0: ready for the next launch)
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    float sum = 0.f;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += 256) sum += __ldcg(partials + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float total = 0.f;
#pragma unroll
        for (int w = 0; w < 256 / 32; ++w) total += s_part[w];
        *loss = total;
    }
}

constexpr unsigned kLossMaxCtas = 148 * 8;

}  // namespace
}  // namespace jsb

namespace jsb {
// partial sums + ticket of the loss kernels of one (device, stream): launches on one stream are
// ordered, launches on different streams get different scratch.  Never freed (4.7 KB each).
static float *loss_scratch(cudaStream_t stream) {
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, float *> pool;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    float *&p = pool[std::make_pair(dev, stream)];
    if (!p) {
        const size_t bytes = (kLossMaxCtas + 4) * sizeof(float);
        if (cudaMalloc((void **)&p, bytes) != cudaSuccess) { p = nullptr; return nullptr; }
        if (cudaMemset(p, 0, bytes) != cudaSuccess) { cudaFree(p); p = nullptr; return nullptr; }
    }
    return p;
}

template <typename TargetT>
static int mse_loss_grad_impl(cudaStream_t stream, const float *img, const TargetT *target, size_t n,
                              float scale, float *v_img, float *loss) {
    if (!loss || (n && (!img || !target || !v_img))) {
        set_error("mse_loss_grad: null buffer");
        return 1;
    }
    if (n == 0) {
        JSB_CUDA_OK(cudaMemsetAsync(loss, 0, sizeof(float), stream));
        return 0;
    }
    float *scratch = loss_scratch(stream);
    if (!scratch) {
        set_error("mse_loss_grad: no scratch for the partial sums");
        return 1;
    }
    // (four elements per thread: 16-byte image / cotangent accesses, 16- or 4-byte target ones)
    const bool vec = ((reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(v_img)) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(target) & (4 * sizeof(TargetT) - 1)) == 0;
    const size_t work = vec ? (n + 3) / 4 : n;
    unsigned blocks = (unsigned)((work + 255) / 256 < kLossMaxCtas ? (work + 255) / 256 : kLossMaxCtas);
    if (blocks == 0) blocks = 1;
    launch_chained(mse_loss_grad_kernel<TargetT>, blocks, 256, 0, stream, img, target, n, scale,
                   1.0f / (float)n, v_img, loss, scratch,
                   reinterpret_cast<unsigned *>(scratch + kLossMaxCtas), vec);
    JSB_CUDA_OK(cudaGetLastError());
    note_launch();
    return 0;
}
}  // namespace jsb

extern "C" int jsb200_mse_loss_grad(jsb200_stream_t stream, const float *img, const float *target,
                                    size_t n, float scale, float *v_img, float *loss) {
    return jsb::mse_loss_grad_impl<float>((cudaStream_t)stream, img, target, n, scale, v_img, loss);
}

extern "C" int jsb200_mse_loss_grad_u8(jsb200_stream_t stream, const float *img,
                                       const unsigned char *target, size_t n, float scale,
                                       float *v_img, float *loss) {
    return jsb::mse_loss_grad_impl<unsigned char>((cudaStream_t)stream, img, target, n, scale, v_img, loss);
}
