// Depth sort of N (u32 key, u32 id) pairs: this library's sort (csrc/radix_sort.cu) against
// cub::DeviceRadixSort::SortPairs<uint32,uint32> on the SAME keys, same box, CUDA events.
// The keys are float bits of depths drawn like the benchmark scenes' (uniform in [lo, hi]), so
// only ~24 bits are live; CUB is timed both on all 32 bits and on the live bit range it would be
// given if the caller knew it (begin_bit..end_bit).
// build (tools/microbench/Makefile): nvcc -gencode arch=compute_100a,code=sm_100a -O3
//        depth_sort_vs_cub.cu ../../jaxsplat_b200/csrc/radix_sort.cu -o depth_sort_vs_cub
// usage: depth_sort_vs_cub [N=1000000] [reps=200]  -> one JSON line
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cub/cub.cuh>

#include "../../jaxsplat_b200/csrc/kernels.cuh"

namespace jsb {
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fputc('\n', stderr);
}
void note_launch(int) {}
}  // namespace jsb

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

static float median(std::vector<float> v) {
    std::sort(v.begin(), v.end());
    return v[v.size() / 2];
}

int main(int argc, char **argv) {
    const unsigned n = argc > 1 ? (unsigned)atol(argv[1]) : 1000000u;
    const int reps = argc > 2 ? atoi(argv[2]) : 200;
    const float lo = 5.0f, hi = 11.0f;
    std::vector<unsigned> hk(n), hv(n);
    unsigned long long s = 88172645463325252ull;
    for (unsigned i = 0; i < n; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        const float d = lo + (hi - lo) * (float)((s >> 11) & 0xFFFFFF) / 16777216.0f;
        memcpy(&hk[i], &d, 4);
        hv[i] = i;
    }
    unsigned *k0, *v0, *ka, *kb, *va, *vb;
    CK(cudaMalloc(&k0, n * 4)); CK(cudaMalloc(&v0, n * 4));
    CK(cudaMalloc(&ka, n * 4)); CK(cudaMalloc(&kb, n * 4));
    CK(cudaMalloc(&va, n * 4)); CK(cudaMalloc(&vb, n * 4));
    CK(cudaMemcpy(k0, hk.data(), n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(v0, hv.data(), n * 4, cudaMemcpyHostToDevice));
    // reference order on the host: stable sort by key
    std::vector<unsigned> order(n);
    for (unsigned i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](unsigned a, unsigned b) { return hk[a] < hk[b]; });

    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    // flush buffer: > L2 so that every timed run starts with cold keys, like inside a render step
    const size_t flush_bytes = 256u << 20;
    void *flush;
    CK(cudaMalloc(&flush, flush_bytes));

    auto check = [&](const unsigned *dv, const char *what) {
        std::vector<unsigned> out(n);
        CK(cudaMemcpy(out.data(), dv, n * 4, cudaMemcpyDeviceToHost));
        size_t bad = 0;
        for (unsigned i = 0; i < n; ++i) bad += out[i] != order[i];
        if (bad) fprintf(stderr, "%s: %zu of %u values out of order\n", what, bad, n);
        return bad == 0;
    };

    // ---- ours ----
    const size_t scratch_bytes = jsb::radix_scratch_bytes(n, 32);
    void *scratch;
    int *err;
    CK(cudaMalloc(&scratch, scratch_bytes));
    CK(cudaMalloc(&err, 16));
    CK(cudaMemset(err, 0, 16));
    std::vector<float> t_ours, t_live, t_cub32, t_cublive;
    bool ok_ours = true, ok_live = true, ok_cub = true;
    for (int live = 0; live < 2; ++live) {
        for (int r = 0; r < reps + 5; ++r) {
            CK(cudaMemcpyAsync(ka, k0, n * 4, cudaMemcpyDeviceToDevice, st));
            CK(cudaMemcpyAsync(va, v0, n * 4, cudaMemcpyDeviceToDevice, st));
            CK(cudaMemsetAsync(flush, r, flush_bytes, st));
            bool in_b = false;
            CK(cudaEventRecord(e0, st));
            int rc = live ? jsb::radix_sort_depth_live(st, ka, kb, va, vb, n, scratch, err, &in_b)
                          : jsb::radix_sort_pairs<uint32_t>(st, ka, kb, va, vb, n, 32, scratch, err, &in_b);
            CK(cudaEventRecord(e1, st));
            CK(cudaStreamSynchronize(st));
            if (rc) return 1;
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r >= 5) (live ? t_live : t_ours).push_back(ms);
            if (r == 0) (live ? ok_live : ok_ours) = check(in_b ? vb : va, live ? "ours(live)" : "ours");
        }
    }
    // ---- CUB ----
    size_t temp_bytes = 0;
    cub::DoubleBuffer<unsigned> dk(ka, kb), dv(va, vb);
    CK(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, dk, dv, (int)n, 0, 32, st));
    void *temp;
    CK(cudaMalloc(&temp, temp_bytes));
    unsigned kmin = 0xffffffffu, kmax = 0;
    for (unsigned i = 0; i < n; ++i) { kmin = std::min(kmin, hk[i]); kmax = std::max(kmax, hk[i]); }
    int end_bit = 32;
    while (end_bit > 0 && ((kmax >> (end_bit - 1)) == (kmin >> (end_bit - 1)))) --end_bit;  // common prefix
    for (int live = 0; live < 2; ++live) {
        for (int r = 0; r < reps + 5; ++r) {
            CK(cudaMemcpyAsync(ka, k0, n * 4, cudaMemcpyDeviceToDevice, st));
            CK(cudaMemcpyAsync(va, v0, n * 4, cudaMemcpyDeviceToDevice, st));
            CK(cudaMemsetAsync(flush, r, flush_bytes, st));
            cub::DoubleBuffer<unsigned> bk(ka, kb), bv(va, vb);
            CK(cudaEventRecord(e0, st));
            CK(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, bk, bv, (int)n, 0, live ? end_bit : 32, st));
            CK(cudaEventRecord(e1, st));
            CK(cudaStreamSynchronize(st));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r >= 5) (live ? t_cublive : t_cub32).push_back(ms);
            if (r == 0) ok_cub = ok_cub && check(bv.Current(), "cub");
        }
    }
    auto mn = [](const std::vector<float> &v) { return *std::min_element(v.begin(), v.end()); };
    printf("{\"n\": %u, \"reps\": %d, \"key_range\": [%.1f, %.1f], \"live_bits\": %d, "
           "\"ours_us\": {\"median\": %.2f, \"min\": %.2f, \"order_ok\": %s}, "
           "\"ours_live_us\": {\"median\": %.2f, \"min\": %.2f, \"order_ok\": %s}, "
           "\"cub_32bit_us\": {\"median\": %.2f, \"min\": %.2f}, "
           "\"cub_live_bits_us\": {\"median\": %.2f, \"min\": %.2f}, \"cub_order_ok\": %s, "
           "\"cub_temp_bytes\": %zu, \"note\": \"cold keys (256 MB memset between runs), CUDA events on the sort's stream\"}\n",
           n, reps, lo, hi, end_bit, 1e3f * median(t_ours), 1e3f * mn(t_ours), ok_ours ? "true" : "false",
           1e3f * median(t_live), 1e3f * mn(t_live), ok_live ? "true" : "false",
           1e3f * median(t_cub32), 1e3f * mn(t_cub32), 1e3f * median(t_cublive), 1e3f * mn(t_cublive),
           ok_cub ? "true" : "false", temp_bytes);
    return (ok_ours && ok_cub) ? 0 : 2;
}
