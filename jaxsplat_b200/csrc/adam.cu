// Adam update for the parameter arrays either side of the hot path: the step of
// the reference's examples/single_image.py:32-40 (optax.adam(lr): b1 0.9, b2 0.999, eps 1e-8, no
// weight decay), as ONE streaming kernel over a flat float array -- 16 B/element read, 12 B/element
// written, 16-byte vectors -- instead of the ~10 elementwise kernels of an unfused update.
//   m <- b1 m + (1-b1) g ; v <- b2 v + (1-b2) g^2 ; p <- p - lr * (m / bc1) / (sqrt(v / bc2) + eps)
// with bc1 = 1 - b1^t, bc2 = 1 - b2^t evaluated on the host in double (optax's bias correction).
#include "../../include/jaxsplat_b200.h"
#include "common.cuh"

namespace jsb {
namespace {

struct AdamParams {
    float lr, b1, b2, eps, inv_bc1, inv_bc2;
};

__device__ __forceinline__ void adam_one(float &p, float g, float &m, float &v, const AdamParams &a) {
    m = fmaf(a.b1, m, (1.f - a.b1) * g);
    v = fmaf(a.b2, v, (1.f - a.b2) * g * g);
    const float mhat = m * a.inv_bc1;
    const float vhat = v * a.inv_bc2;
    p -= a.lr * (mhat / (sqrtf(vhat) + a.eps));
}

__global__ void __launch_bounds__(256)
adam_step_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                 float *__restrict__ v, size_t n, const AdamParams a, bool vec) {
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t done = 0;
    if (vec) {
        const size_t nv = n >> 2;
        for (size_t i = tid; i < nv; i += stride) {
            float4 P = reinterpret_cast<float4 *>(p)[i];
            const float4 G = __ldg(reinterpret_cast<const float4 *>(g) + i);
            float4 M = reinterpret_cast<float4 *>(m)[i];
            float4 V = reinterpret_cast<float4 *>(v)[i];
            adam_one(P.x, G.x, M.x, V.x, a);
            adam_one(P.y, G.y, M.y, V.y, a);
            adam_one(P.z, G.z, M.z, V.z, a);
            adam_one(P.w, G.w, M.w, V.w, a);
            reinterpret_cast<float4 *>(p)[i] = P;
            reinterpret_cast<float4 *>(m)[i] = M;
            reinterpret_cast<float4 *>(v)[i] = V;
        }
        done = nv << 2;
    }
    for (size_t i = done + tid; i < n; i += stride) adam_one(p[i], g[i], m[i], v[i], a);
}

}  // namespace
}  // namespace jsb

extern "C" int jsb200_adam_step(jsb200_stream_t stream_, float *param, const float *grad, float *m,
                                float *v, size_t n, float lr, float b1, float b2, float eps,
                                unsigned long long step) {
    using namespace jsb;
    if (n == 0) return 0;
    if (!param || !grad || !m || !v || step == 0) {
        set_error("adam_step: null buffer or step 0 (steps count from 1)");
        return 1;
    }
    AdamParams a;
    a.lr = lr; a.b1 = b1; a.b2 = b2; a.eps = eps;
    double p1 = 1.0, p2 = 1.0;
    {   // b^step by squaring, in double
        double x1 = b1, x2 = b2;
        for (unsigned long long e = step; e; e >>= 1) {
            if (e & 1) { p1 *= x1; p2 *= x2; }
            x1 *= x1; x2 *= x2;
        }
    }
    a.inv_bc1 = (float)(1.0 / (1.0 - p1));
    a.inv_bc2 = (float)(1.0 / (1.0 - p2));
    const bool vec = ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
                       reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    const size_t work = vec ? (n + 3) / 4 : n;
    unsigned blocks = (unsigned)((work + 255) / 256 < 148 * 8 ? (work + 255) / 256 : 148 * 8);
    if (blocks == 0) blocks = 1;
    adam_step_kernel<<<blocks, 256, 0, (cudaStream_t)stream_>>>(param, grad, m, v, n, a, vec);
    JSB_CUDA_OK(cudaGetLastError());
    note_launch();
    return 0;
}
