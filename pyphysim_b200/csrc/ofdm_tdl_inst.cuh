// ofdm_tdl_inst.cuh — defines launch_ofdm_tdl<B200_NR, B200_NT>; included by one .cu per antenna
// configuration so the (large) kernel instantiations compile in parallel.
#include <type_traits>

#include "ofdm_tdl_fpair.cuh"

namespace b200phy {

#ifndef B200_PAIR_STATIC_SHAPE
#define B200_PAIR_STATIC_SHAPE 1
#endif

template <typename T, bool FUSED, int NR, int NT, bool WSG>
static int launch_one(const OfdmP &p, const Modem &m, const void *table, uint64_t first_unit,
                      int64_t n_units, const uint8_t *idx, const void *phi, const void *psi,
                      const void *noise, uint8_t *idx_hat, void *eq_out, int64_t *counters,
                      size_t smem, cudaStream_t st) {
    auto kern = ofdm_tdl_kernel<T, FUSED, NR, NT, WSG>;
    int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)),
                       "cudaFuncSetAttribute(ofdm_tdl_kernel)");
    if (e) return e;
    int dev = 0, sms = 148, occ = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    e = check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kOT, smem),
                   "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    if (e) return e;
    if (occ < 1) { set_error("ofdm_tdl_kernel does not fit on an SM (%zu B shared memory)", smem); return B200PHY_ERR_UNSUPPORTED; }
    long long grid = (long long)sms * occ;
    if (grid > n_units) grid = n_units;
    cx<T> *ws = nullptr;
    if (WSG) {
        e = check_cuda(cudaMallocAsync((void **)&ws, sizeof(cx<T>) * size_t(grid) * (NR + 1) * p.fft, st),
                       "cudaMallocAsync(workspace)");
        if (e) return e;
    }
    kern<<<int(grid), kOT, smem, st>>>(p, m, (const cx<T> *)table, first_unit, (long long)n_units, idx,
                                       (const T *)phi, (const T *)psi, (const cx<T> *)noise, idx_hat,
                                       (cx<T> *)eq_out, ws, (unsigned long long *)counters);
    count_launch();
    note_kernel("ofdm_tdl_kernel<%s,%d,%d,%d,%d>", sizeof(T) == 4 ? "float" : "double", int(FUSED), NR, NT, int(WSG));
    e = check_cuda(cudaGetLastError(), "ofdm_tdl_kernel launch");
    if (WSG) cudaFreeAsync(ws, st);
    return e;
}

// threads per CTA of the pair kernel for the big shapes (Nr Nt > 4: one CTA per SM, limited by shared
// memory) when fft is a multiple of 2048.  512 threads = 16 warps per SM instead of 8 at 128 registers per thread
// (the 4x4 detection spills ~100 B).  Measured on C5 (r02, same box, three alternating runs): stream mode +0.3 %,
// fused RNG +5.8 % -> 512 is the default (-DB200_PAIR_BIG_KT=256 builds the 8-warp variant).  All 16 warps run
// the same phase between barriers, which is why doubling the warps buys so little: see DESIGN.md 4.2.
#ifndef B200_PAIR_BIG_KT
#define B200_PAIR_BIG_KT 512
#endif

template <bool FUSED, int NR, int NT, bool QAMK, int KT, int LGF, bool TC = false>
static int launch_pair_kt(const OfdmP &p, const Modem &m, const void *table, uint64_t first_unit, int64_t n_units,
                          const uint8_t *idx, const void *phi, const void *psi, const void *noise,
                          uint8_t *idx_hat, void *eq_out, int64_t *counters, size_t smem, cudaStream_t st) {
    auto kern = ofdm_tdl_pair_kernel<FUSED, NR, NT, QAMK, KT, LGF, TC>;
    int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)),
                       "cudaFuncSetAttribute(ofdm_tdl_pair_kernel)");
    if (e) return e;
    int dev = 0, sms = 148, occ = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    e = check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, KT, smem),
                   "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    if (e) return e;
    if (occ < 1) return -1;                       // does not fit: caller falls back to the generic kernel
    if constexpr (TC) {
        // This instantiation allocates tensor memory (tcgen05.alloc, 128 of the SM's 512 columns).  The occupancy API
        // cannot know the column count and answers 1; the real limit is min(shared memory, registers, 512 / 128).
        cudaFuncAttributes fa;
        int smem_sm = 0, regs_sm = 0;
        if (cudaFuncGetAttributes(&fa, kern) == cudaSuccess &&
            cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev) == cudaSuccess) {
            int o = int(size_t(smem_sm) / (smem + fa.sharedSizeBytes + 1024));
            const int by_regs = regs_sm / (fa.numRegs * KT);
            if (by_regs < o) o = by_regs;
            if (o > 4) o = 4;
            if (o > occ) occ = o;
        }
    }
    long long grid = (long long)sms * occ;
    if (grid > n_units) grid = n_units;
    kern<<<int(grid), KT, smem, st>>>(p, m, (const cx<float> *)table, first_unit, (long long)n_units, idx,
                                      (const float *)phi, (const float *)psi, (const cx<float> *)noise,
                                      idx_hat, (cx<float> *)eq_out, (unsigned long long *)counters);
    count_launch();
    note_kernel("ofdm_tdl_pair_kernel<%d,%d,%d,%d,%d,%d,%d>", int(FUSED), NR, NT, int(QAMK), KT, LGF, int(TC));
    return check_cuda(cudaGetLastError(), "ofdm_tdl_pair_kernel launch");
}

// compile-time frame shape of the headline config (full band, fft 1024, 2x2); everything else takes the
// run-time-shape instantiation

template <bool FUSED, int NR, int NT, bool QAMK>
static int launch_pair_k(const OfdmP &p, const Modem &m, const void *table, uint64_t first_unit, int64_t n_units,
                       const uint8_t *idx, const void *phi, const void *psi, const void *noise,
                       uint8_t *idx_hat, void *eq_out, int64_t *counters, size_t smem, cudaStream_t st) {
    constexpr int KTB = (NR * NT > 4) ? B200_PAIR_BIG_KT : kOT;
    if constexpr (KTB != kOT) {
        if ((p.fft & (KTB * kJBC - 1)) == 0)
            return launch_pair_kt<FUSED, NR, NT, QAMK, KTB, 0>(p, m, table, first_unit, n_units, idx, phi, psi, noise, idx_hat, eq_out, counters, smem, st);
    }
#if B200_PAIR_STATIC_SHAPE
    // measured: +12 % on the 2x2 headline; the 4-antenna kernels (1 CTA/SM, 170-220 registers) get slower
    // with the fully unrolled FFT stages (C5 -12 %), so they keep the run-time shape
    if constexpr (NR == 2 && NT == 2) {
        // per-subcarrier channel matrices on the tensor cores (tcgen05, 3xTF32): see ofdm_tdl_pair.cuh
        if (p.fft == 1024 && p.used == p.fft && ofdm_tdl_pair_tc_ok(p))
            return launch_pair_kt<FUSED, NR, NT, QAMK, kOT, 10, true>(p, m, table, first_unit, n_units, idx, phi, psi, noise, idx_hat, eq_out, counters, smem, st);
    }
    if constexpr (NR * NT <= 4) {
        if (p.fft == 1024 && p.used == p.fft)
            return launch_pair_kt<FUSED, NR, NT, QAMK, kOT, 10>(p, m, table, first_unit, n_units, idx, phi, psi, noise, idx_hat, eq_out, counters, smem, st);
    }
#endif
    return launch_pair_kt<FUSED, NR, NT, QAMK, kOT, 0>(p, m, table, first_unit, n_units, idx, phi, psi, noise, idx_hat, eq_out, counters, smem, st);
}

template <bool FUSED, int NR, int NT>
static int launch_pair(const OfdmP &p, const Modem &m, const void *table, uint64_t first_unit, int64_t n_units,
                       const uint8_t *idx, const void *phi, const void *psi, const void *noise,
                       uint8_t *idx_hat, void *eq_out, int64_t *counters, size_t smem, cudaStream_t st) {
    if (m.kind == B200PHY_MODEM_QAM)
        return launch_pair_k<FUSED, NR, NT, true>(p, m, table, first_unit, n_units, idx, phi, psi, noise, idx_hat, eq_out, counters, smem, st);
    return launch_pair_k<FUSED, NR, NT, false>(p, m, table, first_unit, n_units, idx, phi, psi, noise, idx_hat, eq_out, counters, smem, st);
}

template <bool FUSED, bool QAMK, int LGF>
static int launch_fpair_kl(const OfdmP &p, const Modem &m, const void *table, uint64_t first_unit, int64_t n_units,
                        const uint8_t *idx, const void *phi, const void *psi, const void *noise,
                        uint8_t *idx_hat, void *eq_out, int64_t *counters, size_t smem, cudaStream_t st) {
    auto kern = ofdm_tdl_fpair_kernel<FUSED, QAMK, LGF>;
    int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)),
                       "cudaFuncSetAttribute(ofdm_tdl_fpair_kernel)");
    if (e) return e;
    int dev = 0, sms = 148, occ = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    e = check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kOT, smem),
                   "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    if (e) return e;
    if (occ < 1) return -1;
    const int64_t n_pairs = (n_units + 1) / 2;
    long long grid = (long long)sms * occ;
    if (grid > n_pairs) grid = n_pairs;
    kern<<<int(grid), kOT, smem, st>>>(p, m, (const cx<float> *)table, first_unit, (long long)n_units, idx,
                                       (const float *)phi, (const float *)psi, (const cx<float> *)noise,
                                       idx_hat, (cx<float> *)eq_out, (unsigned long long *)counters);
    count_launch();
    note_kernel("ofdm_tdl_fpair_kernel<%d,%d,%d>", int(FUSED), int(QAMK), LGF);
    return check_cuda(cudaGetLastError(), "ofdm_tdl_fpair_kernel launch");
}

template <bool FUSED, bool QAMK>
static int launch_fpair_k(const OfdmP &p, const Modem &m, const void *table, uint64_t first_unit, int64_t n_units,
                        const uint8_t *idx, const void *phi, const void *psi, const void *noise,
                        uint8_t *idx_hat, void *eq_out, int64_t *counters, size_t smem, cudaStream_t st) {
#if B200_PAIR_STATIC_SHAPE
    if (p.fft == 1024 && p.used == p.fft)
        return launch_fpair_kl<FUSED, QAMK, 10>(p, m, table, first_unit, n_units, idx, phi, psi, noise, idx_hat, eq_out, counters, smem, st);
#endif
    return launch_fpair_kl<FUSED, QAMK, 0>(p, m, table, first_unit, n_units, idx, phi, psi, noise, idx_hat, eq_out, counters, smem, st);
}

template <bool FUSED>
static int launch_fpair(const OfdmP &p, const Modem &m, const void *table, uint64_t first_unit, int64_t n_units,
                        const uint8_t *idx, const void *phi, const void *psi, const void *noise,
                        uint8_t *idx_hat, void *eq_out, int64_t *counters, size_t smem, cudaStream_t st) {
    if (m.kind == B200PHY_MODEM_QAM)
        return launch_fpair_k<FUSED, true>(p, m, table, first_unit, n_units, idx, phi, psi, noise, idx_hat, eq_out, counters, smem, st);
    return launch_fpair_k<FUSED, false>(p, m, table, first_unit, n_units, idx, phi, psi, noise, idx_hat, eq_out, counters, smem, st);
}

template <typename T, int NR, int NT>
static int launch_typed(const OfdmP &p, const Modem &m, const void *table, uint64_t first_unit,
                        int64_t n_units, const uint8_t *idx, const void *phi, const void *psi,
                        const void *noise, uint8_t *idx_hat, void *eq_out, int64_t *counters,
                        cudaStream_t st) {
    int dev = 0, max_smem = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if constexpr (std::is_same<T, float>::value && NR == 1 && NT == 1) {
        // SISO: two frames per CTA on the two lanes of the packed FP32 pipe (an odd last frame runs in a
        // final pair with a masked second lane, so every frame takes the same arithmetic whatever the batch)
        if (ofdm_tdl_fpair_ok(p)) {
            const size_t fs = ofdm_tdl_fpair_smem(p, m.M);
            if (fs <= size_t(max_smem)) {
                const int e = idx ? launch_fpair<false>(p, m, table, first_unit, n_units, idx, phi, psi, noise, idx_hat, eq_out, counters, fs, st)
                                  : launch_fpair<true>(p, m, table, first_unit, n_units, idx, phi, psi, noise, idx_hat, eq_out, counters, fs, st);
                if (e >= 0) return e;
            }
        }
    }
    if constexpr (std::is_same<T, float>::value && (NR % 2 == 0) && (NT % 2 == 0)) {
        if (ofdm_tdl_pair_ok(p, NR, NT)) {
            const size_t ps = ofdm_tdl_pair_smem(p, m.M, NR, NT);
            if (ps <= size_t(max_smem)) {
                const int e = idx ? launch_pair<false, NR, NT>(p, m, table, first_unit, n_units, idx, phi, psi, noise, idx_hat, eq_out, counters, ps, st)
                                  : launch_pair<true, NR, NT>(p, m, table, first_unit, n_units, idx, phi, psi, noise, idx_hat, eq_out, counters, ps, st);
                if (e >= 0) return e;
            }
        }
    }
    size_t smem = ofdm_tdl_smem<T>(p, m.M, NR, NT, false);
    const bool wsg = smem > size_t(max_smem);
    if (wsg) {
        smem = ofdm_tdl_smem<T>(p, m.M, NR, NT, true);
        if (smem > size_t(max_smem)) {
            set_error("OFDM/TDL frame needs %zu B of shared memory (> %d): fft/cp/taps too large", smem, max_smem);
            return B200PHY_ERR_UNSUPPORTED;
        }
    }
    const bool fused = !idx;
#define B200_GO(F, W) launch_one<T, F, NR, NT, W>(p, m, table, first_unit, n_units, idx, phi, psi, noise, idx_hat, eq_out, counters, smem, st)
    if (fused) return wsg ? B200_GO(true, true) : B200_GO(true, false);
    return wsg ? B200_GO(false, true) : B200_GO(false, false);
#undef B200_GO
}

template <>
int launch_ofdm_tdl<B200_NR, B200_NT>(int dtype, const OfdmP &p, const Modem &m, const void *table,
                                      uint64_t first_unit, int64_t n_units, const uint8_t *idx,
                                      const void *phi, const void *psi, const void *noise,
                                      uint8_t *idx_hat, void *eq_out, int64_t *counters,
                                      cudaStream_t st) {
    if (dtype == B200PHY_F32)
        return launch_typed<float, B200_NR, B200_NT>(p, m, table, first_unit, n_units, idx, phi, psi, noise, idx_hat, eq_out, counters, st);
    return launch_typed<double, B200_NR, B200_NT>(p, m, table, first_unit, n_units, idx, phi, psi, noise, idx_hat, eq_out, counters, st);
}

}  // namespace b200phy
