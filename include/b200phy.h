/*
 * b200phy.h — C ABI of libb200phy.so: the B200-native (sm_100a) implementation of
 * pyphysim's per-realization link hot path.
 *
 * The reference (darcamo/pyphysim) has no FFI layer: its boundary is the Python
 * class API.  Each entry point below names the reference method(s) it replaces
 * (paths relative to the reference repo root).  The Python façade package
 * `pyphysim_b200` binds these symbols with ctypes; INTEGRATION.md shows the stub a
 * maintainer would add to the reference itself.
 *
 * Conventions
 *   - Every pointer argument marked "dev" is a CUDA device pointer; "host" is a
 *     host pointer.  No C++/torch types cross this boundary.
 *   - `dtype` selects the arithmetic: B200PHY_F32 (complex64 samples, interleaved
 *     re,im floats) or B200PHY_F64 (complex128).  Small linear solves always run in
 *     double.  Symbol indices are int64 in the stage ops (reference dtype) and
 *     uint8 in the fused link ops (M <= 256).
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  All work
 *     is enqueued on it; nothing synchronises unless documented.
 *   - Return value: 0 on success, otherwise a B200PHY_ERR_* code;
 *     b200phy_last_error() gives the message (thread-local).
 *   - counters: int64[4] = {symbol_errors, bit_errors, num_symbols, num_bits},
 *     ACCUMULATED in place (zero them before a new SNR point).
 *   - Random draws follow the shared Philox4x32-10 contract (oracle/philox.py,
 *     pyphysim_b200/csrc/rng.cuh): a pure function of (seed, stream, unit, word),
 *     unit = global realization / frame index, so results are invariant to batch
 *     size and to how units are sharded over GPUs.
 */
#ifndef B200PHY_H
#define B200PHY_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200PHY_VERSION 101

enum { B200PHY_F32 = 0, B200PHY_F64 = 1 };

/* demap kinds */
enum {
    B200PHY_MODEM_TABLE = 0, /* arbitrary table: argmin_m |s_m - r|, first minimum wins
                                (Modulator.demodulate, modulators/fundamental.py:241-248) */
    B200PHY_MODEM_QAM = 1,   /* square Gray QAM built by QAM.__init__ (fundamental.py:659-777):
                                per-axis slicer, identical decisions to the table search */
    B200PHY_MODEM_BPSK = 2,  /* BPSK.modulate/demodulate (fundamental.py:605-647) */
    B200PHY_MODEM_QPSK = 3   /* QPSK() = PSK(4, pi/4) with its Gray order (fundamental.py:396-448, 510-531):
                                symbols[i] lies in the quadrant (re < 0 iff bit 0, im < 0 iff bit 1), so the
                                min-distance search is a quadrant slicer with identical decisions (ties on the
                                axes, a null set, go to the lower index like argmin); map still reads the table */
};

enum {
    B200PHY_OK = 0,
    B200PHY_ERR_INVALID = 1,     /* bad argument (the façade maps these to ValueError) */
    B200PHY_ERR_UNSUPPORTED = 2, /* shape outside what the kernels are built for */
    B200PHY_ERR_CUDA = 3,        /* CUDA runtime error */
    B200PHY_ERR_RANGE = 4        /* symbol index >= M seen on the device (-> ValueError,
                                    fundamental.py:196-199) */
};

#define B200PHY_MAX_TAPS 32
#define B200PHY_MAX_RAYS 64
#define B200PHY_MAX_ANT 4

/* Jakes evaluation strategies of the fused OFDM/TDL kernel */
enum {
    B200PHY_JAKES_AUTO = 0,     /* polynomial when its error bound holds, else recurrence */
    B200PHY_JAKES_RECURRENCE = 1,
    B200PHY_JAKES_POLY = 2
};

typedef struct b200phy_modem {
    int32_t kind;      /* B200PHY_MODEM_* */
    int32_t M;         /* constellation size, power of two, <= 256 */
    const void *table; /* dev: M complex values of the call's dtype (Modulator.symbols) */
} b200phy_modem;

/* One OFDM frame over a Jakes/TDL channel (notebooks/TDL_and_OFDM.ipynb cell 32; MIMO
 * variant = Blast.encode -> OFDM.modulate per tx antenna -> TdlMimoChannel.corrupt_data ->
 * OFDM.demodulate per rx antenna -> per-subcarrier Blast.decode). */
typedef struct b200phy_ofdm_tdl_params {
    int32_t struct_size;  /* sizeof(b200phy_ofdm_tdl_params), for ABI checking */
    int32_t dtype;        /* B200PHY_F32 / B200PHY_F64 */
    int32_t fft, cp, used, n_sym; /* OFDM(fft, cp, used) (modulators/ofdm.py:20-94); OFDM symbols/frame */
    int32_t Nr, Nt;       /* Nt <= Nr <= 4.  1x1: one-tap equaliser (ofdm.py:469-552); else Blast (mimo/mimo.py:465-660) */
    int32_t n_taps;       /* discretised profile (channels/fading.py:272-304) */
    int32_t L;            /* Jakes rays (channels/fading_generators.py:319-351) */
    int32_t jakes_mode;   /* B200PHY_JAKES_* */
    int32_t reserved;     /* flags; bit 0 = do not use the antenna-pair FFMA2 kernel (A/B, tests); bit 1 = compute the
                             per-subcarrier channel matrices of the 2x2 / fft-1024 link on the tensor cores (tcgen05,
                             3xTF32; measured 3 % slower than the CUDA-core form, hence opt-in) */
    int32_t delays[B200PHY_MAX_TAPS];    /* tap delays in samples, increasing */
    double tap_powers[B200PHY_MAX_TAPS]; /* linear tap powers */
    double Fd, Ts, t0;    /* Doppler [Hz], sample time [s], time of the frame's first sample */
    double noise_var;     /* AWGN variance per rx sample (1/dB2Linear(SNR)) */
    double filter_noise_var; /* Blast.set_noise_var value: >0 MMSE, 0 ZF (mimo.py:597-605) */
    uint64_t seed;        /* Philox key */
} b200phy_ofdm_tdl_params;

/* ---- library ------------------------------------------------------------------ */
int b200phy_version(void);
const char *b200phy_last_error(void);
/* number of kernels this library has launched in the calling process (bench `gpu_launches`) */
uint64_t b200phy_launch_count(void);
/* "name<template arguments>" of the fused-link kernel instantiation launched last (process-wide; bench.py
 * checks it against the kernel name in the committed ncu capture before it quotes that capture) */
const char *b200phy_last_kernel(void);

/* ---- stage ops (API-parity path behind the façade classes) ---------------------- */

/* Modulator.modulate (fundamental.py:175-199), BPSK.modulate (:605-630).
 * idx dev int64[n] -> out dev complex[n].  err_flag dev int32[1] is set to 1 if any idx >= M
 * (negative indices wrap like NumPy). */
int b200phy_map(int dtype, const b200phy_modem *modem, const int64_t *idx, int64_t n, void *out,
                int32_t *err_flag, void *stream);

/* Modulator.demodulate (fundamental.py:201-248), BPSK.demodulate (:632-647).
 * r dev complex[n] -> idx_hat dev int64[n]. */
int b200phy_demap(int dtype, const b200phy_modem *modem, const void *r, int64_t n,
                  int64_t *idx_hat, void *stream);

/* sum(a != b) and util.misc.count_bit_errors (util/misc.py:519-566) in one pass.
 * a, b dev int64[n]; out dev int64[2] += {symbol_errors, bit_errors}. */
int b200phy_count_errors(const int64_t *a, const int64_t *b, int64_t n, int64_t *out,
                         void *stream);

/* util.misc.count_bits (util/misc.py:449-476): elementwise popcount of non-negative int64. */
int b200phy_count_bits(const int64_t *a, int64_t n, int64_t *out, void *stream);

/* misc.randn_c * sqrt(noise_var) added in place (util/misc.py:327-355): x[i] += sigma*c_i with c_i
 * the complex normal `first + i` of (seed, stream_id, unit). */
int b200phy_awgn(int dtype, void *x, int64_t n, double noise_var, uint64_t seed,
                 uint32_t stream_id, uint64_t unit, uint64_t first, void *stream);

/* OFDM.modulate (ofdm.py:394-429): x dev complex[batch][n_sym*used] ->
 * out dev complex[batch][n_sym*(fft+cp)].  (Zero padding to whole symbols is done by the caller.) */
int b200phy_ofdm_mod(int dtype, const void *x, void *out, int64_t batch, int n_sym, int fft, int cp,
                     int used, void *stream);
/* OFDM.demodulate (ofdm.py:431-466): r dev complex[batch][n_sym*(fft+cp)] -> y dev complex[batch][n_sym*used]. */
int b200phy_ofdm_demod(int dtype, const void *r, void *y, int64_t batch, int n_sym, int fft, int cp,
                       int used, void *stream);

/* JakesSampleGenerator.generate_more_samples (fading_generators.py:495-523).
 * phi, psi dev real[L][P]; h dev complex[P][N]; t_n = t0 + n*Ts*1.0000000001 (:459-467). */
int b200phy_jakes(int dtype, const void *phi, const void *psi, int L, int64_t P, int64_t N, double Fd,
                  double Ts, double t0, void *h, void *stream);

/* TdlChannel.corrupt_data (channels/fading.py:1046-1124), forward direction.
 * x dev complex[Nt][N]; fading dev complex[n_taps][Nr][Nt][N] (unit-power Jakes/Rayleigh samples);
 * y dev complex[Nr][N+mem], mem = delays[n_taps-1].  tap_powers host double[n_taps] (linear),
 * delays host int32[n_taps]. */
int b200phy_tdl_apply(int dtype, const void *x, const void *fading, const double *tap_powers,
                      const int32_t *delays, int n_taps, int Nr, int Nt, int64_t N, void *y,
                      void *stream);

/* TdlImpulseResponse.get_freq_response (channels/fading.py:513-536): DFT along the delay axis of the
 * zero-padded taps for every time sample, evaluated directly on the sparse taps:
 * out[k][a][n] = sum_l taps[l][a][n] exp(-2 pi i k d_l / fft);  taps dev complex[n_taps][A][N],
 * out dev complex[fft][A][N], delays host int32[n_taps]. */
int b200phy_tdl_freq_response(int dtype, const void *taps, const int32_t *delays, int n_taps, int64_t A,
                              int64_t N, int fft, void *out, void *stream);

/* Block-static frequency-domain channel, TdlChannel.corrupt_data_in_freq_domain (channels/fading.py:
 * 1126-1287): block b (bs consecutive symbols) is multiplied by the frequency response of the b-th
 * impulse response at the used carriers: y[r][b*bs+i] = sum_t H[car[i]][r][t][b] * x[t][b*bs+i].
 * H dev complex[fft][Nr][Nt][B] (b200phy_tdl_freq_response of the per-block taps), x dev complex[Nt][B*bs],
 * carriers dev int32[bs] or NULL (= 0..fft-1, bs == fft), y dev complex[Nr][B*bs]. */
int b200phy_freq_apply(int dtype, const void *H, const void *x, const int32_t *carriers, int fft, int bs,
                       int Nr, int Nt, int64_t B, void *y, void *stream);

/* x[row][col] *= scales[row] in place (tap power scaling fading.py:949-953, path loss
 * singleuser.py:130-151, Blast 1/sqrt(Nt) mimo.py:639-640).  scales host double[rows], rows <= 64. */
int b200phy_scale_rows(int dtype, void *x, int rows, int64_t cols, const double *scales, void *stream);

/* OfdmOneTapEqualizer.equalize_data (ofdm.py:515-552) restated as FFT(mean taps):
 * y dev complex[Nr][n_sym*used]; fading as in b200phy_tdl_apply with N = n_sym*(fft+cp);
 * out dev complex[Nt*n_sym*used]: 1x1 -> y/H; otherwise per-subcarrier Blast decode
 * (filter_noise_var > 0: MMSE, else ZF; mimo.py:264-309, 590-660), layer-interleaved. */
int b200phy_ofdm_equalize(int dtype, const void *y, const void *fading, const double *tap_powers,
                          const int32_t *delays, int n_taps, int Nr, int Nt, int n_sym, int fft,
                          int cp, int used, double filter_noise_var, void *out, void *stream);

/* Blast.decode (mimo.py:642-660) for a batch of channels: H dev complex[batch][Nr][Nt],
 * y dev complex[batch][Nr][T] -> out dev complex[batch][Nt*T] (order='F' interleave). */
int b200phy_blast_decode(int dtype, const void *H, const void *y, int64_t batch, int Nr, int Nt,
                         int T, double filter_noise_var, void *out, void *stream);

/* Alamouti.encode (mimo.py:1166-1214): s dev complex[batch][T] -> x dev complex[batch][2][T]. */
int b200phy_alamouti_encode(int dtype, const void *s, int64_t batch, int T, void *x, void *stream);
/* Alamouti.decode (mimo.py:1216-1287): H dev complex[batch][Nr][2], y dev complex[batch][Nr][T]
 * -> out dev complex[batch][T]. */
int b200phy_alamouti_decode(int dtype, const void *H, const void *y, int64_t batch, int Nr, int T,
                            void *out, void *stream);

/* Thin SVD of a batch of small channel matrices, always in double (what SVDMimo / GMDMimo take from
 * np.linalg.svd, mimo.py:855-898, 974-1019): H dev complex128[batch][Nr][Nt], 1 <= Nt <= Nr <= 4 ->
 * U dev complex128[batch][Nr][Nt], S dev double[batch][Nt] (descending), V dev complex128[batch][Nt][Nt]
 * with H = U diag(S) V^H.  Gauge: the largest-magnitude entry of every column of V is real positive
 * (numpy's gauge is LAPACK's; the two differ by one unit phase per singular pair). */
int b200phy_svd(const void *H, int64_t batch, int Nr, int Nt, void *U, double *S, void *V, void *stream);
/* util.misc.gmd (misc.py:18-159, tol = 0) for a batch: U dev complex128[batch][Nr][Nt], S dev
 * double[batch][Nt], V dev complex128[batch][Nt][Nt] (V, not V^H) -> Q like U, R dev double[batch][Nt][Nt]
 * (upper triangular, constant diagonal), P like V, with U diag(S) V^H = Q R P^H. */
int b200phy_gmd(const void *U, const double *S, const void *V, int64_t batch, int Nr, int Nt, void *Q,
                double *R, void *P, void *stream);
/* Y = A X: A dev complex[rows][cols] (rows, cols <= 8), X dev complex[cols][n] -> Y dev complex[rows][n];
 * the W.dot(X) / G_H.dot(Y) of SVDMimo / GMDMimo / MRT encode and decode (mimo.py:737-783, 900-948,
 * 1021-1067). */
int b200phy_mat_apply(int dtype, const void *A, int rows, int cols, const void *X, int64_t n, void *Y,
                      void *stream);

/* ---- reference signals and pilot-based channel estimation ------------- */

/* One user's reference-signal sequence, generated on the device in double and stored in `dtype`:
 *   out[n] = scale * base[n mod Nzc] * exp(j 2 pi n_cs n / denominator),  n < size
 * base = exp(-j pi u m (m + 1 + 2q) / Nzc) (calcBaseZC, reference_signals/zadoffchu.py:11-36), or, when
 * phi_table (host int8[Nzc], Nzc = 12 or 24: a row of 3GPP TS 36.211 Table 5.5.1.2-1 / -2) is given,
 * base = exp(j pi phi[m] / 4) (RootSequence, root_sequence.py:273-283).  n mod Nzc is the cyclic extension
 * (get_extended_ZF, zadoffchu.py:75-113); the shift is get_shifted_root_seq (:39-72) with denominator 8
 * for SRS (srs.py:23-48) and 12 for DMRS (dmrs.py:19-41); `scale` carries the cover-code entry and the
 * normalisation of UeSequence (srs.py:71-93).  out dev complex[size]. */
int b200phy_refsig_sequence(int dtype, int Nzc, int u, double q_re, double q_im, const int8_t *phi_table,
                            int size, int n_cs, int denominator, double scale_re, double scale_im, void *out,
                            void *stream);

/* CazacBasedChannelEstimator.estimate_channel_freq_domain (reference_signals/channel_estimation.py:69-131)
 * for a batch of received vectors, with the cover-code average of CazacBasedWithOCCChannelEstimator
 * (:163-251) folded in: z = conj(ref) * mean_c(cover[c] * y[c]); taps = ifft(z, Nsc)[0 : num_taps_to_keep+1];
 * out = scale * fft(taps, size_multiplier * Nsc).  ref dev complex[Nsc]; y dev complex[batch][n_cover][Nsc];
 * cover host double[2*n_cover] (re, im) or NULL (= ones), n_cover <= 4; out dev complex[batch][size_multiplier*Nsc]. */
int b200phy_cazac_estimate(int dtype, const void *ref_seq, const void *y, const double *cover_re_im, int n_cover,
                           int64_t batch, int Nsc, int num_taps_to_keep, int size_multiplier, double scale,
                           void *out, void *stream);

/* compute_ls_estimation (channel_estimation/estimators.py:12-61): H = Y s^H (s s^H)^-1 per realization.
 * Y dev complex[batch][Nr][P]; s dev complex[batch][Nt][P] (s_per_unit = 1) or complex[Nt][P] shared by all
 * realizations (s_per_unit = 0); out dev complex[batch][Nr][Nt].  Nt <= 4. */
int b200phy_ls_estimate(int dtype, const void *Y, const void *s, int s_per_unit, int64_t batch, int Nr, int Nt,
                        int P, void *out, void *stream);

/* compute_mmse_estimation (estimators.py:100-174), one transmit antenna:
 * h = (noise_power I + P C)^-1 C (Y s^H) P / (s s^H).  Y, s as above with Nt = 1; C host complex128[Nr][Nr]
 * as (re, im) doubles; out dev complex[batch][Nr].  Nr <= 8. */
int b200phy_mmse_estimate(int dtype, const void *Y, const void *s, int s_per_unit, int64_t batch, int Nr, int P,
                          double noise_power, const double *C_re_im, void *out, void *stream);

/* ---- fused link ops (throughput path) ------------------------------------------
 * Each covers realizations/frames [first_unit, first_unit + n_units).  Draw arrays are
 * "stream mode" inputs; pass them all NULL for "fused mode" (in-kernel Philox).
 * idx_hat / dec_out are optional outputs (NULL to skip).
 * Alignment (vector loads): stream-mode arrays must start on their natural vector boundary — idx / idx_hat
 * 4 bytes (2 for Alamouti), complex / real draw arrays 16 bytes for the flat SISO and Alamouti links and one
 * element for the others — i.e. pass whole contiguous tensors, not slices at odd offsets; a misaligned
 * pointer returns B200PHY_ERR_INVALID. */

/* 1 symbol per realization: r = h*x + sqrt(noise_var)*n; r /= h
 * (notebooks/Transmission_with_Rayleigh_and_AWGN_channels.ipynb cell 8; rayleigh=0 is the AWGN
 * chain of apps/awgn_modulators/simulate_psk.py:51-115).
 * idx dev uint8[n], h dev complex[n] (ignored if !rayleigh), noise dev complex[n]. */
int b200phy_link_siso_flat(int dtype, const b200phy_modem *modem, int rayleigh, double noise_var,
                           uint64_t seed, uint64_t first_unit, int64_t n_units, const uint8_t *idx,
                           const void *h, const void *noise, uint8_t *idx_hat, void *dec_out,
                           int64_t *counters, void *stream);

/* Alamouti over flat Rayleigh (apps/mimo/simulate_mimo.py:68-142 with mimo.Alamouti):
 * per realization H[Nr][2], S symbols (even).  idx dev uint8[n][S], H dev complex[n][Nr][2],
 * noise dev complex[n][Nr][S]. */
int b200phy_link_alamouti(int dtype, const b200phy_modem *modem, int Nr, int S, double noise_var,
                          uint64_t seed, uint64_t first_unit, int64_t n_units, const uint8_t *idx,
                          const void *H, const void *noise, uint8_t *idx_hat, void *dec_out,
                          int64_t *counters, void *stream);

/* Blast ZF/MMSE over flat Rayleigh (simulate_mimo.py:68-142 with mimo.Blast): S symbol vectors per
 * realization.  idx dev uint8[n][S*Nt], H dev complex[n][Nr][Nt], noise dev complex[n][Nr][S]. */
int b200phy_link_blast(int dtype, const b200phy_modem *modem, int Nr, int Nt, int S, double noise_var,
                       double filter_noise_var, uint64_t seed, uint64_t first_unit, int64_t n_units,
                       const uint8_t *idx, const void *H, const void *noise, uint8_t *idx_hat,
                       void *dec_out, int64_t *counters, void *stream);

/* Channel-dependent precoding over flat Rayleigh (simulate_mimo.py:68-142 with mimo.SVDMimo,
 * mimo.GMDMimo or mimo.MRT; mimo.py:666-783, 829-1067).  SVD / GMD: square Nr == Nt in [2, 4], Nt layers,
 * symbol p = l*S + s of a realization is layer l at time s (X = transmit_data.reshape(Nt, -1));
 * filter_noise_var > 0 selects the MMSE filter of the GMD equivalent channel, else ZF (ignored by SVD
 * and MRT).  MRT: Nr == 1, one layer.  idx dev uint8[n][S*layers], H dev complex[n][Nr][Nt],
 * noise dev complex[n][Nr][S]. */
#define B200PHY_MIMO_SVD 1
#define B200PHY_MIMO_GMD 2
#define B200PHY_MIMO_MRT 3
int b200phy_link_precoded(int dtype, const b200phy_modem *modem, int scheme, int Nr, int Nt, int S,
                          double noise_var, double filter_noise_var, uint64_t seed, uint64_t first_unit,
                          int64_t n_units, const uint8_t *idx, const void *H, const void *noise,
                          uint8_t *idx_hat, void *dec_out, int64_t *counters, void *stream);

/* OFDM over Jakes/TDL, SISO or Blast-MIMO (see b200phy_ofdm_tdl_params).
 * idx dev uint8[n][Nt*n_sym*used]; phi, psi dev real[n][L][n_taps][Nr][Nt];
 * noise dev complex[n][Nr][N+mem] (unit variance), N = n_sym*(fft+cp);
 * idx_hat dev uint8 like idx; eq_out dev complex[n][Nt*n_sym*used] (equalised symbols);
 * rx_out dev complex[n][Nr][n_sym*used]: what OFDM.demodulate returns per rx antenna (ofdm.py:431-466),
 * i.e. the samples BEFORE the equaliser / Blast.decode — the quantity the 1e-5 sample tolerance is held on. */
int b200phy_link_ofdm_tdl(const b200phy_ofdm_tdl_params *p, const b200phy_modem *modem,
                          uint64_t first_unit, int64_t n_units, const uint8_t *idx, const void *phi,
                          const void *psi, const void *noise, uint8_t *idx_hat, void *eq_out,
                          void *rx_out, int64_t *counters, void *stream);

/* Validates a parameter block exactly as b200phy_link_ofdm_tdl would (struct size, dtype, the OFDM.set_parameters
 * checks of modulators/ofdm.py:75-90, tap / ray / antenna ranges, Jakes mode) without touching the device. */
int b200phy_ofdm_tdl_check_params(const b200phy_ofdm_tdl_params *p);

/* ---- draw dumps: write the fused-mode Philox draws in the stream-mode layouts, so the oracle can
 * consume literally the numbers the device used.  NULL outputs are skipped. */
int b200phy_draw_siso_flat(int dtype, int bits, uint64_t seed, uint64_t first_unit, int64_t n_units,
                           uint8_t *idx, void *h, void *noise, void *stream);
int b200phy_draw_flat_mimo(int dtype, int bits, int Nr, int Nt, int S, int n_data, uint64_t seed,
                           uint64_t first_unit, int64_t n_units, uint8_t *idx, void *H, void *noise,
                           void *stream);
int b200phy_draw_ofdm_tdl(const b200phy_ofdm_tdl_params *p, int bits, uint64_t first_unit,
                          int64_t n_units, uint8_t *idx, void *phi, void *psi, void *noise,
                          void *stream);

/* ---- host-buffer entry points (what a non-CUDA caller binds; `e2e` in bench.py) -------------
 * Same semantics as the device versions but every array is a HOST pointer (pinned memory makes the
 * copies asynchronous), and the constellation is passed as M (re, im) double pairs.
 *   Monte Carlo mode — all draw arrays and idx_hat NULL: what one `_run_simulation` call of a
 *   SimulationRunner needs (simulations/runner.py:1491-1517 merges only the counters).  One kernel over all
 *   units; the parameters go in, 32 bytes of counters come back.
 *   Stream mode — draw arrays and / or idx_hat given: the library stages ~64 MiB chunks through its own
 *   device buffers, overlapping H2D, compute and D2H on two internal streams.
 * Returns after the results are in host memory.  counters is host int64[4], accumulated.  Arguments are
 * validated before anything is copied; on error nothing is left in flight. */
int b200phy_link_siso_flat_host(int dtype, int modem_kind, int M, const double *table_re_im,
                                int rayleigh, double noise_var, uint64_t seed, uint64_t first_unit,
                                int64_t n_units, const uint8_t *idx, const void *h,
                                const void *noise, uint8_t *idx_hat, int64_t *counters);
/* simulate_mimo.py:68-142 with mimo.Alamouti / mimo.Blast / mimo.SVDMimo, GMDMimo, MRT (the device versions above) */
int b200phy_link_alamouti_host(int dtype, int modem_kind, int M, const double *table_re_im, int Nr, int S,
                               double noise_var, uint64_t seed, uint64_t first_unit, int64_t n_units,
                               const uint8_t *idx, const void *H, const void *noise, uint8_t *idx_hat,
                               int64_t *counters);
int b200phy_link_blast_host(int dtype, int modem_kind, int M, const double *table_re_im, int Nr, int Nt, int S,
                            double noise_var, double filter_noise_var, uint64_t seed, uint64_t first_unit,
                            int64_t n_units, const uint8_t *idx, const void *H, const void *noise,
                            uint8_t *idx_hat, int64_t *counters);
int b200phy_link_precoded_host(int dtype, int modem_kind, int M, const double *table_re_im, int scheme, int Nr,
                               int Nt, int S, double noise_var, double filter_noise_var, uint64_t seed,
                               uint64_t first_unit, int64_t n_units, const uint8_t *idx, const void *H,
                               const void *noise, uint8_t *idx_hat, int64_t *counters);
int b200phy_link_ofdm_tdl_host(const b200phy_ofdm_tdl_params *p, int modem_kind, int M,
                               const double *table_re_im, uint64_t first_unit, int64_t n_units,
                               const uint8_t *idx, const void *phi, const void *psi,
                               const void *noise, uint8_t *idx_hat, int64_t *counters);

#ifdef __cplusplus
}
#endif
#endif /* B200PHY_H */
