"""Device-side diagnostics of the fused links (no oracle involved: product kernel against product kernel).

`precision_drift` answers the question DESIGN.md §6 leaves open for the throughput arithmetic: the reference
computes in complex128, the benchmarked kernels in complex64 — how far do the error COUNTERS of a float32
run drift from the float64 run on literally the same draws, and how close to a decision boundary is every
symbol the two disagree on?  The float64 kernels are the ones held decision-exact to the NumPy oracle
(tests/test_gpu_ofdm_tdl.py), so this transfers that anchor to full-size batches in seconds."""
import numpy as np

from . import _lib


def _margins(table, z):
    """Gap between the distances to the nearest and second-nearest constellation point, on the device."""
    torch = _lib.torch_cuda()
    d = (table.reshape(-1, 1) - z.reshape(1, -1)).abs()
    two = torch.topk(d, 2, dim=0, largest=False).values
    return two[1] - two[0]


def precision_drift(link, n_units, first_unit=0, chunk=4096):
    """Run `link` (an OfdmTdlLink) in float32 and in float64 over units [first_unit, first_unit + n_units) on the
    SAME draws (the float32 draws of the Philox stream, widened exactly to float64) and compare decisions.

    Returns a dict: units, symbols, symbol_errors_f32 / _f64, bit_errors_f32 / _f64, symbol_count_drift,
    bit_count_drift (f32 - f64), decision_mismatches, mismatch_rate, worst_mismatch_margin (largest float64
    decision margin among the symbols the two arithmetics decide differently; 0 when they never differ),
    median_margin (of all symbols, for scale)."""
    torch = _lib.torch_cuda()
    l32, l64 = link.with_dtype('f32'), link.with_dtype('f64')
    table = torch.from_numpy(np.asarray(link.modulator.symbols, dtype=np.complex128)).cuda()
    acc32 = torch.zeros(4, dtype=torch.int64, device='cuda')
    acc64 = torch.zeros(4, dtype=torch.int64, device='cuda')
    mism, worst, med = 0, 0.0, []
    for off in range(0, n_units, chunk):
        n = min(chunk, n_units - off)
        idx, phi, psi, noise = l32.draw(first_unit + off, n)
        _, hat32 = l32.run(n, first_unit=first_unit + off, draws=(idx, phi, psi, noise), counters=acc32,
                           want_idx=True)
        d64 = (idx, phi.double(), psi.double(), noise.to(torch.complex128))
        del phi, psi, noise
        _, hat64, eq64 = l64.run(n, first_unit=first_unit + off, draws=d64, counters=acc64, want_idx=True,
                                 want_eq=True)
        bad = (hat32 != hat64).reshape(-1).nonzero().reshape(-1)
        mism += int(bad.numel())
        if bad.numel():
            worst = max(worst, float(_margins(table, eq64.reshape(-1)[bad]).max()))
        if off == 0:
            med.append(float(_margins(table, eq64.reshape(-1)[:1 << 16]).median()))
        del d64, hat32, hat64, eq64
    c32, c64 = acc32.cpu().numpy(), acc64.cpu().numpy()
    return {"units": int(n_units), "symbols": int(c64[2]),
            "symbol_errors_f32": int(c32[0]), "symbol_errors_f64": int(c64[0]),
            "bit_errors_f32": int(c32[1]), "bit_errors_f64": int(c64[1]),
            "symbol_count_drift": int(c32[0] - c64[0]), "bit_count_drift": int(c32[1] - c64[1]),
            "decision_mismatches": mism, "mismatch_rate": mism / max(1, int(c64[2])),
            "worst_mismatch_margin": worst, "median_margin": med[0] if med else None}
