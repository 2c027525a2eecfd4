// Mean-squared-error loss and its gradient in ONE streaming kernel (the metric is
// fwd+bwd of loss = mean((img - target)^2), the loss of the reference's
// examples/single_image.py:21-30 and examples/benchmark.py).  An eager host framework spends
// three kernels and 150 MB of traffic on it at 1920x1080 (subtract, dot, scale); this reads the
// image and the target once and writes the cotangent once (75 MB):
//   v_img[i] = (img[i] - target[i]) * scale          (scale = 2 / n from the caller, one rounding
//                                                      each: the eager result bit for bit)
//   loss    += sum_i (img[i] - target[i])^2 / n       (per-CTA partial sums, one atomic per CTA;
//                                                      `loss` is cleared by a 4-byte memset)
#include "../../include/jaxsplat_b200.h"
#include "common.cuh"

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
                     bool vec) {
    __shared__ float s_part[256 / 32];
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
        atomicAdd(loss, total * inv_n);
    }
}

}  // namespace
}  // namespace jsb

namespace jsb {
template <typename TargetT>
static int mse_loss_grad_impl(cudaStream_t stream, const float *img, const TargetT *target, size_t n,
                              float scale, float *v_img, float *loss) {
    if (!loss || (n && (!img || !target || !v_img))) {
        set_error("mse_loss_grad: null buffer");
        return 1;
    }
    JSB_CUDA_OK(cudaMemsetAsync(loss, 0, sizeof(float), stream));
    if (n == 0) return 0;
    // (four elements per thread: 16-byte image / cotangent accesses, 16- or 4-byte target ones)
    const bool vec = ((reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(v_img)) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(target) & (4 * sizeof(TargetT) - 1)) == 0;
    const size_t work = vec ? (n + 3) / 4 : n;
    unsigned blocks = (unsigned)((work + 255) / 256 < 148 * 8 ? (work + 255) / 256 : 148 * 8);
    if (blocks == 0) blocks = 1;
    mse_loss_grad_kernel<TargetT><<<blocks, 256, 0, stream>>>(img, target, n, scale, 1.0f / (float)n,
                                                              v_img, loss, vec);
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
