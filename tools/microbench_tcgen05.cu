// microbench_tcgen05.cu — the per-subcarrier channel matrices H[k] = sum_l g_l W^(k d_l) of the 2x2 headline
// (1024 bins x 15 sparse taps x 4 complex entries) as ONE tcgen05 tile per frame.
//
//   W^((k0 + off) d) = W^(k0 d) W^(off d), off = 128 v + 256 u (8 offsets): the bins k0 + off, k0 < 128, share the DFT
//   operand of the bins k0 < 128, with the coefficients rotated by the 8th roots W^(off d).  So
//       D[128, 64] = A[128, 32] B[32, 64]      A[k0][(l, re|im)] = (cos, sin) of W^(k0 d_l)            (constant)
//                                              B[(l, re|im)][(off, entry, re|im)] = rotated coefficients (per frame)
//   3xTF32: D = A_lo B_hi + A_hi B_lo + A_hi B_hi, A_hi / A_lo resident in TMEM (64 columns), D in TMEM (64 columns),
//   B_hi / B_lo in shared memory (K-major, no swizzle, 2 x 8 KB), 12 tcgen05.mma (M = 128, N = 64, K = 8) per frame
//   issued by one thread, completion on an mbarrier, results read back with tcgen05.ld: thread tid gets the bins
//   tid + 256 u, u < 4 — exactly the bins a detection thread of ofdm_tdl_pair.cuh owns.
//
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o build_mb/mb_tcgen05 tools/microbench_tcgen05.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int kFft = 1024, kTaps = 15, kEnt = 4;      // complex entries per tap (Nr x Nt = 2 x 2)
constexpr int kM = 128, kN = 64, kK = 32;             // GEMM tile
__constant__ int c_delay[16];

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo) {
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    hi = __uint_as_float(h);
    lo = x - hi;
}
__device__ __forceinline__ void mbar_init(void *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// K-major, no-swizzle shared-memory operand descriptor: 8-row x 16-byte core matrices; `lbo` = byte distance between
// the two 16-byte K chunks of an MMA, `sbo` = byte distance between 8-row groups
__device__ __forceinline__ uint64_t smem_desc(const void *p, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= uint64_t((smem_u32(p) >> 4) & 0x3fff);
    d |= uint64_t((lbo >> 4) & 0x3fff) << 16;
    d |= uint64_t((sbo >> 4) & 0x3fff) << 32;
    d |= uint64_t(1) << 46;                               // descriptor version (Blackwell)
    return d;                                             // base offset 0, layout type 0 = no swizzle
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, dense
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(kN >> 3) << 17) | (uint32_t(kM >> 4) << 24);

__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(kIdesc), "r"(accumulate), "r"(0u) : "memory");
}

__global__ void __launch_bounds__(256, 3)
hk_tcgen05_kernel(const float2 *__restrict__ coef, float2 *__restrict__ out, float *chk, int n_frames, int *err_flag) {
    __shared__ __align__(128) float bop[2][4][8][2][8][4];   // [hi|lo][k-step][n-group][k-chunk][row][4 k] = 2 x 8 KB
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float2 rot[8][16];                            // W^(off d_l): [offset index 2 u + v ... see below][tap]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // rotation table: column block r = 4 v + u  <->  bins k0 + 128 v + 256 u
    if (tid < 8 * 16) {
        const int r = tid >> 4, l = tid & 15, off = 128 * (r >> 2) + 256 * (r & 3);
        float s, c;
        sincospif(-2.0f * float((off * c_delay[l]) & (kFft - 1)) / float(kFft), &s, &c);
        rot[r][l] = make_float2(c, s);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t t_ahi = tmem, t_alo = tmem + 32, t_d = tmem + 64;

    // A operand into TMEM: warps 0..3, thread = row k0 = 32 warp + lane, 32 columns (l, re|im) for hi and lo
    if (warp < 4) {
        const int k0 = 32 * warp + lane;
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int l = 0; l < 16; ++l) {
            float s = 0.f, c = 0.f;
            if (l < kTaps) sincospif(-2.0f * float((k0 * c_delay[l]) & (kFft - 1)) / float(kFft), &s, &c);
            float h0, l0, h1, l1;
            split_tf32(c, h0, l0);
            split_tf32(s, h1, l1);
            hi[2 * l] = __float_as_uint(h0); hi[2 * l + 1] = __float_as_uint(h1);
            lo[2 * l] = __float_as_uint(l0); lo[2 * l + 1] = __float_as_uint(l1);
        }
        const uint32_t lane_base = uint32_t(32 * warp) << 16;
#define ST32(ADDR, V) asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" \
        ::"r"(ADDR), "r"(V[0]), "r"(V[1]), "r"(V[2]), "r"(V[3]), "r"(V[4]), "r"(V[5]), "r"(V[6]), "r"(V[7]), "r"(V[8]), "r"(V[9]), "r"(V[10]), "r"(V[11]), "r"(V[12]), "r"(V[13]), "r"(V[14]), "r"(V[15]), \
          "r"(V[16]), "r"(V[17]), "r"(V[18]), "r"(V[19]), "r"(V[20]), "r"(V[21]), "r"(V[22]), "r"(V[23]), "r"(V[24]), "r"(V[25]), "r"(V[26]), "r"(V[27]), "r"(V[28]), "r"(V[29]), "r"(V[30]), "r"(V[31]) : "memory")
        ST32(t_ahi + lane_base, hi);
        ST32(t_alo + lane_base, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    float acc = 0.f;
    unsigned parity = 0;
    for (int f = blockIdx.x; f < n_frames; f += gridDim.x) {
        // ---- B operand: element (kk = 2 l + p, n = 8 r + 2 e + q): p = 0 -> (g'.re, g'.im)[q], p = 1 -> (-g'.im, g'.re)[q]
        for (int i = tid; i < 8 * 16 * kEnt; i += 256) {
            const int e = i & 3, l = (i >> 2) & 15, r = i >> 6;
            float2 g = make_float2(0.f, 0.f);
            if (l < kTaps) {
                const float2 c = coef[(size_t(f) * kTaps + l) * kEnt + e], w = rot[r][l];
                g = make_float2(c.x * w.x - c.y * w.y, c.x * w.y + c.y * w.x);
            }
            // rows n0 = 8 r + 2 e (re), n0 + 1 (im); k = 2 l, 2 l + 1
            const int ks = l >> 2, ch = (l >> 1) & 1, kq = (2 * l) & 3;      // k-step, 16-byte chunk, position in the chunk
            const int grp = r, row = 2 * e;                                   // n / 8, n % 8
            float h, lw;
            split_tf32(g.x, h, lw);                                           // (k = 2l, n re) and (k = 2l+1, n im)
            bop[0][ks][grp][ch][row][kq] = h;      bop[1][ks][grp][ch][row][kq] = lw;
            bop[0][ks][grp][ch][row + 1][kq + 1] = h; bop[1][ks][grp][ch][row + 1][kq + 1] = lw;
            split_tf32(g.y, h, lw);                                           // (k = 2l, n im) = g.im ; (k = 2l+1, n re) = -g.im
            bop[0][ks][grp][ch][row + 1][kq] = h;  bop[1][ks][grp][ch][row + 1][kq] = lw;
            bop[0][ks][grp][ch][row][kq + 1] = -h; bop[1][ks][grp][ch][row][kq + 1] = -lw;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t accum = 0;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint64_t bhi = smem_desc(&bop[0][ks][0][0][0][0], 128, 256), blo = smem_desc(&bop[1][ks][0][0][0][0], 128, 256);
                mma_tf32_ts(t_d, t_alo + 8 * ks, bhi, accum); accum = 1;
                mma_tf32_ts(t_d, t_ahi + 8 * ks, blo, 1);
                mma_tf32_ts(t_d, t_ahi + 8 * ks, bhi, 1);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
        }
        // ---- consumers: thread tid owns bins tid + 256 u = k0 + 128 v + 256 u with k0 = tid & 127, v = tid >> 7
        mbar_wait(&mbar, parity);
        parity ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t v[32];
        const uint32_t taddr = t_d + (uint32_t(32 * (warp & 3)) << 16) + 32 * (warp >> 2);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                       "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int e = 0; e < kEnt; ++e) {
                const float re = __uint_as_float(v[8 * u + 2 * e]), im = __uint_as_float(v[8 * u + 2 * e + 1]);
                if (f == 0) out[(tid + 256 * u) * kEnt + e] = make_float2(re, im);
                acc += re + im;
            }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();           // every thread has read D and B may be rebuilt
    }
    chk[blockIdx.x * 256 + tid] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
    (void)err_flag;
}

static const int h_delay[16] = {0, 3, 8, 10, 13, 19, 20, 21, 23, 25, 28, 29, 30, 31, 33, 0};

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount, n_frames = sms * 3 * 24;
    std::vector<float2> coef(size_t(n_frames) * kTaps * kEnt);
    uint32_t st = 4242u;
    auto rnd = [&] { st = st * 1664525u + 1013904223u; return float(int32_t(st)) * (1.0f / 2147483648.0f); };
    for (auto &v : coef) v = make_float2(rnd(), rnd());
    std::vector<double> ref(size_t(kFft) * kEnt * 2);
    double rms = 0;
    for (int k = 0; k < kFft; ++k)
        for (int e = 0; e < kEnt; ++e) {
            double re = 0, im = 0;
            for (int l = 0; l < kTaps; ++l) {
                const double a = -2.0 * M_PI * double((k * h_delay[l]) % kFft) / kFft, wr = cos(a), wi = sin(a);
                const float2 c = coef[size_t(l) * kEnt + e];
                re += c.x * wr - c.y * wi;
                im += c.x * wi + c.y * wr;
            }
            ref[(size_t(k) * kEnt + e) * 2] = re; ref[(size_t(k) * kEnt + e) * 2 + 1] = im;
            rms += re * re + im * im;
        }
    rms = sqrt(rms / (kFft * kEnt));
    float2 *d_coef, *d_out; float *d_chk; int *d_err;
    CK(cudaMalloc(&d_coef, coef.size() * sizeof(float2)));
    CK(cudaMemcpy(d_coef, coef.data(), coef.size() * sizeof(float2), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_out, size_t(kFft) * kEnt * sizeof(float2)));
    CK(cudaMemset(d_out, 0, size_t(kFft) * kEnt * sizeof(float2)));
    CK(cudaMalloc(&d_chk, size_t(sms) * 3 * 256 * 4));
    CK(cudaMalloc(&d_err, 4));
    CK(cudaMemcpyToSymbol(c_delay, h_delay, sizeof(h_delay)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        hk_tcgen05_kernel<<<sms * 3, 256>>>(d_coef, d_out, d_chk, n_frames, d_err);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    std::vector<float2> got(size_t(kFft) * kEnt);
    CK(cudaMemcpy(got.data(), d_out, got.size() * sizeof(float2), cudaMemcpyDeviceToHost));
    double maxerr = 0; int worst = 0;
    for (size_t i = 0; i < got.size(); ++i) {
        const double e = fmax(fabs(got[i].x - ref[2 * i]), fabs(got[i].y - ref[2 * i + 1]));
        if (e > maxerr) { maxerr = e; worst = int(i); }
    }
    printf("{\"experiment\": \"hk_tcgen05\", \"variant\": \"tcgen05_tf32_x3_A_in_tmem\", \"frames\": %d, \"ms\": %.4f, \"frames_per_s\": %.4g, "
           "\"us_sm_per_frame\": %.3f, \"max_err_over_rms_vs_f64\": %.3g, \"worst_index\": %d}\n",
           n_frames, best, n_frames / best * 1e3, best * 1e3 * sms / n_frames, maxerr / rms, worst);
    if (maxerr / rms > 1e-3)
        for (int i = 0; i < 8; ++i)
            printf("  bin %d entry %d: got (%.6f, %.6f) ref (%.6f, %.6f)\n", (worst / kEnt) + 0, i % kEnt, got[(worst / kEnt) * kEnt + i % kEnt].x,
                   got[(worst / kEnt) * kEnt + i % kEnt].y, ref[2 * ((worst / kEnt) * kEnt + i % kEnt)], ref[2 * ((worst / kEnt) * kEnt + i % kEnt) + 1]);
    return 0;
}
