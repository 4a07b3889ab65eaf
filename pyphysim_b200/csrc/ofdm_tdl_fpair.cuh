// ofdm_tdl_fpair.cuh — FFMA2 variant of the fused OFDM/TDL link for SISO (Nr = Nt = 1), float, in the
// slow-fading regime: one CTA simulates TWO consecutive frames at once, one per lane of Blackwell's packed
// FP32 instructions.  A pair sample is a float4 (A.re, B.re, A.im, B.im) for frames (A, B) = (2i, 2i+1):
// the IFFT / FFT of both frames are one Stockham transform on float4 (ofdm_tdl_pair.cuh), the FIR multiplies
// lane-wise (coefficient pairs x sample pairs), the one-tap equaliser runs per lane.  Semantics, draws and
// error bounds are those of ofdm_tdl.cuh.  An odd last frame runs in lane 0 of a final pair whose lane 1 is a
// masked copy of it (same reads, nothing counted or written): the two lanes are independent IEEE operations of
// the same instruction stream, so a frame's result does not depend on its lane, its partner, the batch size or
// the shard boundaries — counters stay a pure function of (seed, unit).
#pragma once
#include "ofdm_tdl_pair.cuh"

namespace b200phy {

// LGF: compile-time log2(fft) of a full-band frame (see ofdm_tdl_pair_kernel), 0 = run-time shape
template <bool FUSED, bool QAMK, int LGF = 0>
__global__ void __launch_bounds__(kOT, 3)
ofdm_tdl_fpair_kernel(const __grid_constant__ OfdmP p, const Modem m_in, const cx<float> *__restrict__ tab_g,
                      uint64_t first_unit, long long n_units, const uint8_t *__restrict__ idx_g,
                      const float *__restrict__ phi_g, const float *__restrict__ psi_g,
                      const cx<float> *__restrict__ noise_g, uint8_t *__restrict__ idx_hat,
                      cx<float> *__restrict__ eq_out, unsigned long long *counters) {
    using T = float;
    const long long n_pairs = (n_units + 1) >> 1;
    Modem m = m_in;
    if (QAMK) m.kind = B200PHY_MODEM_QAM;      // compile-time kind (see ofdm_tdl_pair_kernel)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int fft = LGF ? (1 << LGF) : p.fft, lg = LGF ? LGF : p.lg;
    const int used = LGF ? fft : p.used, half = LGF ? (fft >> 1) : p.half;
    const int S = p.S, mem = p.mem, cp = p.cp;

    unsigned char *sp = smem_raw;
    auto take = [&](size_t bytes) { unsigned char *r = sp; sp += (bytes + 15) & ~size_t(15); return r; };
    cx<T> *tw = (cx<T> *)take(sizeof(cx<T>) * (fft + kTwc));
    float4 *E2 = (float4 *)take(sizeof(float4) * (mem + S));            // [tail | cp | body], lanes = frames
    float4 *body = E2 + mem + cp;
    take(32);                                                           // TMA landing pad (see ofdm_tdl_pair.cuh)
    float4 *pool = (float4 *)take(sizeof(float4) * 2 * fft);            // rx buffer + scratch
    u64 *gbar = (u64 *)take(sizeof(u64) * p.n_taps * 2);                // [tap][re|im], lanes = frames
    float4 *tails = (float4 *)take(p.n_sym > 1 ? sizeof(float4) * mem : 0);
    u64 *coef = (u64 *)take(sizeof(u64) * p.n_taps * 4 * 2);            // [tap][order][re|im]
    cx<T> *tab = (cx<T> *)take(sizeof(cx<T>) * m.M);
    uint8_t *dsym = (uint8_t *)take(2 * used);                        // [frame lane][used]
    T *ph_phi = (T *)take(sizeof(T) * 2 * p.P4);                        // [frame lane][P4]
    T *ph_psi = (T *)take(sizeof(T) * 2 * p.P4);

    for (int i = tid; i < fft; i += kOT) {
        double s, c;
        sincospi(-2.0 * double(i) / double(fft), &s, &c);
        tw[i] = {T(c), T(s)};
    }
    fill_compact_twiddles(tw, fft);
    if (m.kind != B200PHY_MODEM_BPSK)
        for (int k = tid; k < m.M; k += kOT) tab[k] = tab_g[k];
    __shared__ __align__(8) unsigned long long mbar[2];      // [0] noise rows of a symbol, [1] phases of the next pair
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    unsigned sym_err = 0, bit_err = 0;
    const T sigma = T(p.sigma), tx_scale = T(p.tx_scale), rx_scale = T(p.rx_scale);
    const int n_items = p.n_taps * 2;                // (tap, frame lane)
    int G = 1;
    while (G < 16 && (G * 2) * n_items <= kOT) G *= 2;
    const int sub = tid & (G - 1);
    const int ostride = p.n_taps;
    const double wts = p.w0 * p.Ts1, wt0 = p.w0 * p.t0;
    const bool in_w = p.ifft_in_w != 0;
    const bool o3 = p.porder == 3;
    // rx FFT stage 0 straight from the FIR accumulators (one output block per frame)
    const bool rx0_fused = (kJBC == 4) && fft == kOT * kJBC;

    // stream-mode input pipeline, as in ofdm_tdl_pair.cuh: next pair's phases by LDGSTS during the FIR, its
    // data symbols in two registers, raw noise rows into the (unused) rx buffer at frame start
    const bool pf = !FUSED && p.n_sym == 1 && (p.n_data & 3) == 0 && 2 * p.n_data <= 8 * kOT &&
                    (reinterpret_cast<uintptr_t>(idx_g) & 7) == 0;
    const bool pf16 = pf && (p.P & 3) == 0 && aligned16(phi_g) && aligned16(psi_g);
    const bool apipe = !FUSED;
    // TMA input pipeline (see ofdm_tdl_pair.cuh): the noise rows of the two frames as two bulk copies into the rx
    // buffer, the phases of the next pair as four more; one thread issues, completion counted on an mbarrier
    // (cp >= 1 or a 16-byte aligned base: the copy of a row that starts 8 bytes off begins one element before it)
    const bool tma = !FUSED && rx0_fused && mem >= 1 && p.n_sym == 1 && (cp >= 1 || aligned16(noise_g));
    const bool tma_ph = tma && pf16;
    unsigned par_noise = 0, par_phase = 0;
    const cx<T> *nrow0 = nullptr, *nrow1 = nullptr;
    uint2 idx_pre = make_uint2(0u, 0u);
    auto prefetch = [&](long long f) {               // f = first frame of the pair
        const bool ghost = f + 1 >= n_units;         // odd tail: lane 1 re-reads frame f
        if (tma_ph && tid == 0) mbar_expect_tx(&mbar[1], 4u * unsigned(p.P) * sizeof(T));
#pragma unroll
        for (int ln = 0; ln < 2; ++ln) {
            const long long fl = (ln && !ghost) ? f + 1 : f;
            const T *gp = phi_g + size_t(fl) * p.P, *gq = psi_g + size_t(fl) * p.P;
            T *dp = ph_phi + ln * p.P4, *dq = ph_psi + ln * p.P4;
            if (tma_ph) {
                if (tid == 0) {
                    bulk_g2s(dp, gp, unsigned(p.P) * sizeof(T), &mbar[1]);
                    bulk_g2s(dq, gq, unsigned(p.P) * sizeof(T), &mbar[1]);
                }
            } else if (pf16) {
                for (int i = tid; i < (p.P >> 2); i += kOT) {
                    cp_async<16>(dp + 4 * i, gp + 4 * i);
                    cp_async<16>(dq + 4 * i, gq + 4 * i);
                }
            } else {
                for (int i = tid; i < p.P; i += kOT) {
                    cp_async<4>(dp + i, gp + i);
                    cp_async<4>(dq + i, gq + i);
                }
            }
        }
        if (tid < (p.n_data >> 2)) {
            if (!ghost) {
                idx_pre = __ldg(reinterpret_cast<const uint2 *>(idx_g + size_t(f) * p.n_data) + tid);
            } else {
                const uint32_t *w = reinterpret_cast<const uint32_t *>(idx_g + size_t(f) * p.n_data);
                const int nw = p.n_data >> 2, w0 = 2 * tid, w1 = 2 * tid + 1;
                idx_pre = make_uint2(__ldg(w + (w0 < nw ? w0 : w0 - nw)), __ldg(w + (w1 < nw ? w1 : w1 - nw)));
            }
        }
    };
    if constexpr (!FUSED) {
        if (pf && blockIdx.x < n_pairs) prefetch(2 * (long long)blockIdx.x);
        cp_async_commit();
    }

    for (long long pr = blockIdx.x; pr < n_pairs; pr += gridDim.x) {
        const long long frame = 2 * pr;               // lane 0 = frame, lane 1 = frame + 1 (or a masked copy of frame)
        const bool ghost = frame + 1 >= n_units;
        const long long frameB = ghost ? frame : frame + 1;
        float4 *Y = pool, *W = pool + fft;

        // ---- phases of both frames -> shared memory
        if constexpr (FUSED) {
            for (int it = tid; it < 2 * (p.P4 >> 2); it += kOT) {
                const int ln = it / (p.P4 >> 2), b = it - ln * (p.P4 >> 2);
                const uint64_t unit = first_unit + uint64_t(ln ? frameB : frame);
                const uint4 b1 = rng_block(p.seed, STREAM_CHANNEL, unit, uint64_t(b));
                const uint4 b2 = rng_block(p.seed, STREAM_CHANNEL, unit, uint64_t((p.P4 >> 2) + b));
                reinterpret_cast<float4 *>(ph_phi + ln * p.P4)[b] = make_float4(phase_from_word<T>(b1.x), phase_from_word<T>(b1.y),
                                                                                phase_from_word<T>(b1.z), phase_from_word<T>(b1.w));
                reinterpret_cast<float4 *>(ph_psi + ln * p.P4)[b] = make_float4(phase_from_word<T>(b2.x), phase_from_word<T>(b2.y),
                                                                                phase_from_word<T>(b2.z), phase_from_word<T>(b2.w));
            }
        } else if (pf) {
            if (tma_ph) { mbar_wait(&mbar[1], par_phase); par_phase ^= 1u; }
            else cp_async_wait<0>();                 // prefetched during the previous pair
            if (tid < (p.n_data >> 2)) reinterpret_cast<uint2 *>(dsym)[tid] = idx_pre;
        } else {
#pragma unroll
            for (int ln = 0; ln < 2; ++ln)
                for (int i = tid; i < p.P; i += kOT) {
                    ph_phi[ln * p.P4 + i] = __ldg(phi_g + size_t(ln ? frameB : frame) * p.P + i);
                    ph_psi[ln * p.P4 + i] = __ldg(psi_g + size_t(ln ? frameB : frame) * p.P + i);
                }
        }

        for (int s = 0; s < p.n_sym; ++s) {
            const int n_s = s * S;
            // ---------------- P0: data symbols of both frames, noise into Y
            {
                const int w0 = s * used, cnt = used;
                if constexpr (FUSED) {
                    const int b0 = w0 >> 2, nb = ((w0 + cnt - 1) >> 2) - b0 + 1;
                    for (int it = tid; it < 2 * nb; it += kOT) {
                        const int ln = it / nb, b = b0 + it - ln * nb;
                        const uint4 blk = rng_block(p.seed, STREAM_DATA, first_unit + uint64_t(ln ? frameB : frame), uint64_t(b));
#pragma unroll
                        for (int l = 0; l < 4; ++l) {
                            const int w = 4 * b + l - w0;
                            if (w >= 0 && w < cnt) dsym[ln * used + w] = uint8_t(lane_of(blk, l) >> (32 - m.bits));
                        }
                    }
                } else if (!pf) {
#pragma unroll
                    for (int ln = 0; ln < 2; ++ln) {
                        const uint8_t *src = idx_g + size_t(ln ? frameB : frame) * p.n_data + w0;
                        for (int i = tid; i < cnt; i += kOT) dsym[ln * used + i] = src[i];
                    }
                }
                const int m0 = n_s + cp;
                if constexpr (FUSED) {
                    // one thread draws sample pair (j, j + 1) of BOTH frames: two independent Philox chains per
                    // iteration and whole pair samples (one 16-byte store each) instead of scalar stores
                    const int pr0 = m0 >> 1, npr = ((m0 + fft - 1) >> 1) - pr0 + 1;
                    const uint64_t uA = first_unit + uint64_t(frame), uB = first_unit + uint64_t(frameB);
                    for (int i = tid; i < npr; i += kOT) {
                        const int q = pr0 + i, j = 2 * q - m0;
                        const uint4 b0 = rng_block(p.seed, STREAM_NOISE, uA, uint64_t(q));
                        const uint4 b1 = rng_block(p.seed, STREAM_NOISE, uB, uint64_t(q));
                        // unit-variance normals (what the draw kernel stores); sigma is applied by the FIR epilogue's FFMA2
                        if (j >= 0 && j < fft) {
                            const cx<T> c0 = cnormal<T>(b0.x, b0.y), c1 = cnormal<T>(b1.x, b1.y);
                            Y[j] = make_float4(c0.re, c1.re, c0.im, c1.im);
                        }
                        if (j + 1 >= 0 && j + 1 < fft) {
                            const cx<T> c0 = cnormal<T>(b0.z, b0.w), c1 = cnormal<T>(b1.z, b1.w);
                            Y[j + 1] = make_float4(c0.re, c1.re, c0.im, c1.im);
                        }
                    }
                } else if (tma) {
                    // the rows of frame A / frame B as two bulk copies, back to back from 32 bytes before the rx buffer
                    // (a row that starts 8 bytes off a 16-byte boundary is copied from one element earlier)
                    const size_t rowlen = size_t(p.N + mem);
                    const cx<T> *s0 = noise_g + size_t(frame) * rowlen + m0, *s1 = noise_g + size_t(frameB) * rowlen + m0;
                    const int sh0 = int((reinterpret_cast<uintptr_t>(s0) >> 3) & 1), sh1 = int((reinterpret_cast<uintptr_t>(s1) >> 3) & 1);
                    const int c0 = (fft + sh0 + 1) & ~1, c1 = (fft + sh1 + 1) & ~1;
                    cx<T> *land = reinterpret_cast<cx<T> *>(Y) - 4;
                    nrow0 = land + sh0;
                    nrow1 = land + c0 + sh1;
                    if (tid == 0) {
                        fence_proxy_async();
                        mbar_expect_tx(&mbar[0], unsigned(c0 + c1) * sizeof(cx<T>));
                        bulk_g2s(land, s0 - sh0, unsigned(c0) * sizeof(cx<T>), &mbar[0]);
                        bulk_g2s(land + c0, s1 - sh1, unsigned(c1) * sizeof(cx<T>), &mbar[0]);
                    }
                } else if (apipe) {
                    const size_t rowlen = size_t(p.N + mem);
                    const cx<T> *s0 = noise_g + size_t(frame) * rowlen + m0, *s1 = noise_g + size_t(frameB) * rowlen + m0;
                    // sample j of frame A / frame B lands in the two halves of float4 slot j (see ofdm_tdl_pair.cuh):
                    // in-place pair relayout in the FIR epilogue, each thread consumes only its own copies
                    cx<T> *raw = reinterpret_cast<cx<T> *>(Y);
                    for (int j = tid; j < fft; j += kOT) {
                        cp_async<8>(raw + 2 * j, s0 + j);
                        cp_async<8>(raw + 2 * j + 1, s1 + j);
                    }
                    cp_async_commit();
                } else {
                    const size_t rowlen = size_t(p.N + mem);
                    const cx<T> *s0 = noise_g + size_t(frame) * rowlen + m0, *s1 = noise_g + size_t(frameB) * rowlen + m0;
                    for (int j0 = tid; j0 < fft; j0 += 4 * kOT) {
                        cx<T> v0[4], v1[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (j0 + u * kOT < fft) { v0[u] = load_stream(s0 + j0 + u * kOT); v1[u] = load_stream(s1 + j0 + u * kOT); }
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (j0 + u * kOT < fft)
                                Y[j0 + u * kOT] = make_float4(sigma * v0[u].re, sigma * v1[u].re, sigma * v0[u].im, sigma * v1[u].im);
                    }
                }
            }
            __syncthreads();

            // ---------------- A: map + scatter (lanes = frames), fused with IFFT stage 0: the thread that maps bins
            // j + i fft/4 owns butterfly j of the first (twiddle-free) radix-4 stage (as in ofdm_tdl_pair_kernel)
            float4 *in = in_w ? W : body;
            float4 *other = in_w ? body : W;
            for (int j = tid; j < (fft >> 2); j += kOT) {
                ps v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int q = pos_of(j + i * (fft >> 2), fft, used, half);
                    v[i] = {0ull, 0ull};
                    if (q >= 0) {
                        const cx<T> s0 = map_symbol<T>(m, tab, dsym[q]);
                        const cx<T> s1 = map_symbol<T>(m, tab, dsym[used + q]);
                        v[i] = {pk2(tx_scale * s0.re, tx_scale * s1.re), pk2(tx_scale * s0.im, tx_scale * s1.im)};
                    }
                }
                ps y0, y1, y2, y3;
                bfly4<true>(v[0], v[1], v[2], v[3], y0, y1, y2, y3);
                fft_store_stage0(other, j, true, y0, y1, y2, y3);
            }
            for (int i = tid; i < mem; i += kOT)
                E2[i] = (s > 0) ? tails[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            // ---------------- C: ray setup, items (tap, frame lane)
            for (int it0 = 0; it0 < n_items; it0 += kOT / G) {
                const int it = it0 + tid / G;
                const bool act = it < n_items;
                const int l = act ? it >> 1 : 0, ln = it & 1;
                const T amp = T(p.amp[l]);
                cx<T> a0 = {0.f, 0.f}, a1 = a0, a2 = a0, a3 = a0;
                const double cseg = double(n_s + cp - p.delays[l]) + 0.5 * double(fft - 1);
                if (act && p.cos_f32) {
                    const T *pphi = ph_phi + ln * p.P4 + l + sub * ostride;
                    const T *ppsi = ph_psi + ln * p.P4 + l + sub * ostride;
                    cx<T> a[4] = {a0, a0, a0, a0};
                    const float eps = float(fma(wts, cseg, wt0)), wtsf = float(wts);
                    if (o3) ray_moments_f32<true>(a, pphi, ppsi, sub, p.L, G, ostride, eps, wtsf);
                    else ray_moments_f32<false>(a, pphi, ppsi, sub, p.L, G, ostride, eps, wtsf);
                    a0 = amp * a[0]; a1 = amp * a[1]; a2 = amp * a[2]; a3 = amp * a[3];
                } else if (act) {
                    const T *pphi = ph_phi + ln * p.P4 + l + sub * ostride;
                    const T *ppsi = ph_psi + ln * p.P4 + l + sub * ostride;
                    for (int o = sub; o < p.L; o += G, pphi += G * ostride, ppsi += G * ostride) {
                        const double cphi = p.cos_f32 ? double(cosf(*pphi)) : cos(double(*pphi));
                        const double dl = wts * cphi;
                        T sn, cs;
                        cis_phase<T>(fma(dl, cseg, fma(wt0, cphi, double(*ppsi))), &sn, &cs);
                        const cx<T> e = {amp * cs, amp * sn};
                        const T d1 = T(dl), d2 = -0.5f * d1 * d1, d3 = (-1.0f / 3.0f) * d1 * d2;
                        a0.re += e.re;        a0.im += e.im;
                        a1.re -= d1 * e.im;   a1.im += d1 * e.re;
                        a2.re += d2 * e.re;   a2.im += d2 * e.im;
                        a3.re += d3 * e.im;   a3.im -= d3 * e.re;
                    }
                }
                for (int o = G >> 1; o > 0; o >>= 1) {
                    a0.re += __shfl_xor_sync(0xffffffffu, a0.re, o); a0.im += __shfl_xor_sync(0xffffffffu, a0.im, o);
                    a1.re += __shfl_xor_sync(0xffffffffu, a1.re, o); a1.im += __shfl_xor_sync(0xffffffffu, a1.im, o);
                    a2.re += __shfl_xor_sync(0xffffffffu, a2.re, o); a2.im += __shfl_xor_sync(0xffffffffu, a2.im, o);
                    a3.re += __shfl_xor_sync(0xffffffffu, a3.re, o); a3.im += __shfl_xor_sync(0xffffffffu, a3.im, o);
                }
                if (act && sub == 0) {
                    if (p.porder != 3) a3 = {0.f, 0.f};
                    T *cq = reinterpret_cast<T *>(coef) + (l * 4) * 4 + ln;     // [tap][order][re|im][lane]
                    cq[0] = a0.re; cq[2] = a0.im; cq[4] = a1.re; cq[6] = a1.im;
                    cq[8] = a2.re; cq[10] = a2.im; cq[12] = a3.re; cq[14] = a3.im;
                    const T m1 = T(p.mu[0][l]), m2 = T(p.mu[1][l]), m3 = T(p.mu[2][l]);
                    T *gq = reinterpret_cast<T *>(gbar) + p.cls_pos[l] * 4 + ln;      // class-sorted (see G)
                    gq[0] = a0.re + m1 * a1.re + m2 * a2.re + m3 * a3.re;
                    gq[2] = a0.im + m1 * a1.im + m2 * a2.im + m3 * a3.im;
                }
            }
            // ---------------- B: remaining IFFT passes (end in E2.body); the last one also writes the cyclic prefix
            fft_stockham_pair<true, LGF ? kOT : 0>(other, in, tw, fft, lg, 1, false, cp);
            if constexpr (!FUSED) {                   // ray setup done: the phase buffers are free
                if (pf && pr + gridDim.x < n_pairs) prefetch(2 * (pr + gridDim.x));
                cp_async_commit();
            }

            // ---------------- D: FIR, lane-wise: y += g(tau) * x with all four partial products packed
            {
                const float tau0 = float(tid) - 0.5f * float(fft - 1);
                const float4 *xb = E2 + mem + cp + tid;
                for (int jo0 = 0; jo0 < fft; jo0 += kOT * kJBC) {
                    u64 aRR[kJBC], aII[kJBC], aRI[kJBC], aIR[kJBC], tt2[kJBC];
#pragma unroll
                    for (int jb = 0; jb < kJBC; ++jb) {
                        const float tau = tau0 + float(jo0 + jb * kOT);
                        tt2[jb] = pk2(tau, tau);
                        aRR[jb] = aII[jb] = aRI[jb] = aIR[jb] = 0ull;
                    }
                    auto taps = [&](auto order3) {          // specialised on the polynomial order
                        constexpr int NO = decltype(order3)::value ? 4 : 3;
#pragma unroll 2
                        for (int l = 0; l < p.n_taps; ++l) {
                            const float4 *xl = xb + (jo0 - p.delays[l]);
                            u64 cR[NO], cI[NO];
#pragma unroll
                            for (int o = 0; o < NO; ++o) {
                                const ulonglong2 c = reinterpret_cast<const ulonglong2 *>(coef)[l * 4 + o];
                                cR[o] = c.x; cI[o] = c.y;
                            }
#pragma unroll
                            for (int jb = 0; jb < kJBC; ++jb) {
                                const float4 x = xl[jb * kOT];
                                const u64 xR = pk2(x.x, x.y), xI = pk2(x.z, x.w);
                                u64 gR = cR[NO - 1], gI = cI[NO - 1];
#pragma unroll
                                for (int o = NO - 2; o >= 0; --o) {
                                    gR = fma2(gR, tt2[jb], cR[o]);
                                    gI = fma2(gI, tt2[jb], cI[o]);
                                }
                                aRR[jb] = fma2(gR, xR, aRR[jb]);
                                aII[jb] = fma2(gI, xI, aII[jb]);
                                aRI[jb] = fma2(gR, xI, aRI[jb]);
                                aIR[jb] = fma2(gI, xR, aIR[jb]);
                            }
                        }
                    };
                    if (o3) taps(std::true_type{}); else taps(std::false_type{});
                    ps yv[kJBC];
                    if (!FUSED && tma) {
                        mbar_wait(&mbar[0], par_noise);      // both rows have landed
                        par_noise ^= 1u;
                        const u64 sg = pk2(sigma, sigma);
#pragma unroll
                        for (int jb = 0; jb < kJBC; ++jb) {
                            const cx<T> n0 = nrow0[tid + jo0 + jb * kOT], n1 = nrow1[tid + jo0 + jb * kOT];
                            yv[jb].re = fma2(pk2(n0.re, n1.re), sg, sub2(aRR[jb], aII[jb]));
                            yv[jb].im = fma2(pk2(n0.im, n1.im), sg, add2(aRI[jb], aIR[jb]));
                        }
                    } else if (!FUSED && apipe) {
                        // this thread's raw noise slots have landed: y = sigma * noise + FIR, re-laid as pairs
                        cp_async_wait<1>();
                        const u64 sg = pk2(sigma, sigma);
#pragma unroll
                        for (int jb = 0; jb < kJBC; ++jb) {
                            const float4 v = Y[tid + jo0 + jb * kOT];   // (nA.re, nA.im, nB.re, nB.im)
                            // one fused multiply-add, as in every other mode: bit-identical results
                            yv[jb].re = fma2(pk2(v.x, v.z), sg, sub2(aRR[jb], aII[jb]));
                            yv[jb].im = fma2(pk2(v.y, v.w), sg, add2(aRI[jb], aIR[jb]));
                        }
                    } else {
                        // fused RNG: the buffer holds this symbol's unit-variance normals (lanes = frames)
                        const u64 sg = pk2(sigma, sigma);
#pragma unroll
                        for (int jb = 0; jb < kJBC; ++jb) {
                            const ps y = ld_ps(Y + tid + jo0 + jb * kOT);
                            yv[jb].re = fma2(y.re, sg, sub2(aRR[jb], aII[jb]));
                            yv[jb].im = fma2(y.im, sg, add2(aRI[jb], aIR[jb]));
                        }
                    }
                    if (rx0_fused) {
                        // one output block per frame: this thread holds the four inputs tid + i fft/4 of butterfly
                        // `tid` of the rx FFT's first stage -> straight into W
                        if constexpr (kJBC == 4) {
                            ps y0, y1, y2, y3;
                            bfly4<false>(yv[0], yv[1], yv[2], yv[3], y0, y1, y2, y3);
                            fft_store_stage0(W, tid, true, y0, y1, y2, y3);
                        }
                    } else {
#pragma unroll
                        for (int jb = 0; jb < kJBC; ++jb) st_ps(Y + tid + jo0 + jb * kOT, yv[jb]);
                    }
                }
            }
            __syncthreads();
            if (p.n_sym > 1) {
                for (int i = tid; i < mem; i += kOT) tails[i] = E2[S + i];
                __syncthreads();
            }

            // ---------------- F: paired FFT of both frames; when the last pass yields exactly the 4 bins of a detection
            // thread it is left to the detection phase (fft_last_pass, from registers)
            const bool fuse_last = fft_last_fusable(lg, 4);
            {
                float4 *res = rx0_fused ? fft_stockham_pair<false, LGF ? kOT : 0>(W, Y, tw, fft, lg, 1, fuse_last)
                                        : fft_stockham_pair<false, LGF ? kOT : 0>(Y, W, tw, fft, lg, 0, fuse_last);
                if (res != Y) { W = Y; Y = res; }
            }

            // ---------------- G: H_k (both frames packed), one-tap equaliser, demap, count.  The taps are summed per
            // residue class of their delay mod 4 (class-sorted runs OfdmP::cls_*, order 0, 2, 1, 3) and the class
            // sums combined by a 4-point DFT, as in ofdm_tdl_pair_kernel.
            const int kstride = fft >> 2;
            for (int k0 = tid; k0 < kstride; k0 += kOT) {
                ps Sc[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) Sc[c] = {0ull, 0ull};
                auto tap_sum = [&](ps &acc, int j0, int j1) {
                    for (int j = j0; j < j1; ++j) {
                        const cx<T> w = tw[(k0 * p.cls_delay[j]) & (fft - 1)];
                        const u64 WR = pk2(w.re, w.re), WI = pk2(w.im, w.im), NWI = pk2(-w.im, -w.im);
                        const ps g = {gbar[j * 2], gbar[j * 2 + 1]};
                        acc.re = fma2(g.im, NWI, fma2(g.re, WR, acc.re));
                        acc.im = fma2(g.im, WR, fma2(g.re, WI, acc.im));
                    }
                };
                tap_sum(Sc[0], p.cls_start[0], p.cls_start[1]);
                tap_sum(Sc[2], p.cls_start[1], p.cls_start[2]);
                tap_sum(Sc[1], p.cls_start[2], p.cls_start[3]);
                tap_sum(Sc[3], p.cls_start[3], p.cls_start[4]);
                ps Hk[4];
                {
                    const ps a0 = Sc[0] + Sc[2], a1 = Sc[0] - Sc[2], a2 = Sc[1] + Sc[3], a3 = Sc[1] - Sc[3];
                    Hk[0] = a0 + a2;
                    Hk[1] = {add2(a1.re, a3.im), sub2(a1.im, a3.re)};     // a1 - j a3
                    Hk[2] = a0 - a2;
                    Hk[3] = {sub2(a1.re, a3.im), add2(a1.im, a3.re)};     // a1 + j a3
                }
                ps Yv[4];
                if (fuse_last) {
                    fft_last_pass<4>(Y, tw, fft, lg, k0, Yv);
                } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) Yv[u] = ld_ps(Y + k0 + u * kstride);
                }
                // Two copies of the 8 unrolled detections (4 bins x 2 frames): the Monte Carlo path asks for no
                // per-symbol output, and without the stores and their 64-bit index arithmetic its code is a third
                // shorter (the capture showed 17 % instruction-fetch stalls in this phase)
                auto detect = [&](auto out_tag) {
                    constexpr bool OUT = decltype(out_tag)::value;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int k = k0 + u * kstride;
                        const int q = pos_of(k, fft, used, half);
                        if (q < 0) continue;
                        float yr0, yr1, yi0, yi1, hr0, hr1, hi0, hi1;
                        upk2(Yv[u].re, yr0, yr1);
                        upk2(Yv[u].im, yi0, yi1);
                        upk2(Hk[u].re, hr0, hr1);
                        upk2(Hk[u].im, hi0, hi1);
#pragma unroll
                        for (int ln = 0; ln < 2; ++ln) {
                            if (ln && ghost) continue;
                            const cx<T> y = ln ? mk<T>(rx_scale * yr1, rx_scale * yi1) : mk<T>(rx_scale * yr0, rx_scale * yi0);
                            const cx<T> H = ln ? mk<T>(hr1, hi1) : mk<T>(hr0, hi0);
                            const cx<T> z = cdiv(y, H);
                            const int a = dsym[ln * used + q];
                            const int e = demap_symbol<T>(m, tab, z);
                            sym_err += (e != a);
                            bit_err += __popc(e ^ a);
                            if constexpr (OUT) {
                                const size_t o = size_t(frame + ln) * p.n_data + size_t(s * used + q);
                                if (idx_hat) idx_hat[o] = uint8_t(e);
                                if (eq_out) eq_out[o] = z;
                                if (p.rx_out) static_cast<cx<T> *>(p.rx_out)[o] = y;
                            }
                        }
                    }
                };
                if (idx_hat || eq_out || p.rx_out) detect(std::true_type{}); else detect(std::false_type{});
            }
            if (tma) fence_proxy_async();        // this pair's ordinary stores before the next pair's bulk copies
            __syncthreads();
        }   // OFDM symbols
    }       // frame pairs
    flush_counters(sym_err, bit_err, counters);
    if (blockIdx.x == 0 && tid == 0) {
        atomicAdd(&counters[2], (unsigned long long)n_units * p.n_data);
        atomicAdd(&counters[3], (unsigned long long)n_units * p.n_data * m.bits);
    }
}

inline size_t ofdm_tdl_fpair_smem(const OfdmP &p, int M) {
    auto al = [](size_t b) { return (b + 15) & ~size_t(15); };
    size_t s = 0;
    s += al(sizeof(cx<float>) * (p.fft + kTwc));
    s += al(sizeof(float4) * (p.mem + p.S));
    s += 32 + al(sizeof(float4) * 2 * p.fft);
    s += al(sizeof(u64) * p.n_taps * 2);
    s += al(p.n_sym > 1 ? sizeof(float4) * p.mem : 0);
    s += al(sizeof(u64) * p.n_taps * 4 * 2);
    s += al(sizeof(cx<float>) * M);
    s += al(size_t(2) * p.used);
    s += 2 * al(sizeof(float) * 2 * p.P4);
    return s;
}

inline bool ofdm_tdl_fpair_ok(const OfdmP &p) {
    return p.poly && p.nseg == 1 && (p.fft & (kOT * kJBC - 1)) == 0 && p.gbar_poly && !p.no_pair;
}

}  // namespace b200phy
