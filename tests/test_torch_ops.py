"""torch.ops.b200phy.*: the fused links as registered PyTorch operators.
CPU: the ops exist with tensor/scalar-only schemas and refuse CPU tensors.  GPU: every op reproduces the counters of
the `links` wrapper (same C entry point underneath) in Monte Carlo mode and in stream mode."""
import numpy as np
import pytest
import torch

import pyphysim_b200.torch_ops as T
from pyphysim_b200 import _lib


def test_ops_are_registered_with_tensor_scalar_schemas():
    for name in T.OPS:
        op = getattr(torch.ops.b200phy, name)
        schema = op.default._schema
        kinds = {str(a.type) for a in schema.arguments}
        assert kinds <= {'Tensor', 'Optional[Tensor]', 'int', 'float', 'bool', 'List[int]', 'List[float]'}, (name, kinds)
        assert any(a.name == 'counters' and a.alias_info is not None and a.alias_info.is_write
                   for a in schema.arguments), name                      # in-place counters, declared as such
        assert len(schema.returns) == 0


def test_ops_refuse_cpu_tensors():
    table = torch.zeros(4, dtype=torch.complex64)
    cnt = torch.zeros(4, dtype=torch.int64)
    with pytest.raises(RuntimeError):
        torch.ops.b200phy.link_siso_flat(table, _lib.MODEM_QAM, True, 0.1, 1, 0, 16, None, None, None, cnt)


@pytest.mark.gpu
def test_ops_match_links_wrappers():
    from pyphysim_b200 import links
    from pyphysim_b200.modulators import QAM, QPSK
    import bench
    dev = 'cuda'
    seed, n = 1234, 20000

    def cnt():
        return torch.zeros(4, dtype=torch.int64, device=dev)

    q64 = QAM(64)
    t64 = torch.as_tensor(np.asarray(q64.symbols, dtype=np.complex64), device=dev)
    # siso flat: Monte Carlo mode and stream mode (draws of the same units)
    want = links.link_siso_flat(q64, 0.02, n, seed=seed, first_unit=7)
    c = cnt()
    torch.ops.b200phy.link_siso_flat(t64, q64._kind, True, 0.02, seed, 7, n, None, None, None, c)
    assert c.cpu().tolist() == want.tolist()
    idx, h, noise = links.draw_siso_flat(q64, n, seed=seed, first_unit=7)
    c2, hat = cnt(), torch.empty(n, dtype=torch.uint8, device=dev)
    torch.ops.b200phy.link_siso_flat(t64, q64._kind, True, 0.02, seed, 7, n, idx, h, noise, c2, hat)
    assert c2.cpu().tolist() == want.tolist()
    assert int((hat != idx).sum()) == int(want[0])
    # Alamouti (C4 shape)
    qp = QPSK()
    tq = torch.as_tensor(np.asarray(qp.symbols, dtype=np.complex64), device=dev)
    want = links.link_alamouti(qp, 0.1, n, Nr=2, num_symbols=2, seed=seed)
    c = cnt()
    torch.ops.b200phy.link_alamouti(tq, qp._kind, 2, 2, 0.1, seed, 0, n, None, None, None, c)
    assert c.cpu().tolist() == want.tolist()
    # Blast MMSE 4x4 and SVD precoding
    q16 = QAM(16)
    t16 = torch.as_tensor(np.asarray(q16.symbols, dtype=np.complex64), device=dev)
    want = links.link_blast(q16, 0.01, n, Nr=4, Nt=4, num_symbols=2, filter_noise_var=0.01, seed=seed)
    c = cnt()
    torch.ops.b200phy.link_blast(t16, q16._kind, 4, 4, 2, 0.01, 0.01, seed, 0, n, None, None, None, c)
    assert c.cpu().tolist() == want.tolist()
    want = links.link_precoded(q16, 0.01, n, scheme='svd', Nr=4, Nt=4, num_symbols=2, seed=seed)
    c = cnt()
    torch.ops.b200phy.link_precoded(t16, q16._kind, links.MIMO_SCHEMES['svd'], 4, 4, 2, 0.01, 0.0, seed, 0, n, None,
                                    None, None, c)
    assert c.cpu().tolist() == want.tolist()
    # the headline OFDM / TDL link, f32 and f64 (table dtype selects the arithmetic)
    w = bench.WORKLOADS['ofdm1024_qam64_mimo2x2_tdl']
    link = bench.make_link(w)
    p = link.params
    frames = 64
    for dtype, cplx in (('f32', np.complex64), ('f64', np.complex128)):
        lk = link.with_dtype(dtype)
        want = lk.run(frames, first_unit=3)
        tab = torch.as_tensor(np.asarray(lk.modulator.symbols, dtype=cplx), device=dev)
        c = cnt()
        torch.ops.b200phy.link_ofdm_tdl(tab, lk.modulator._kind, p.fft, p.cp, p.used, p.n_sym, p.Nr, p.Nt,
                                        [int(p.delays[i]) for i in range(p.n_taps)],
                                        [float(p.tap_powers[i]) for i in range(p.n_taps)], p.Fd, p.Ts, p.t0, p.L,
                                        p.noise_var, p.filter_noise_var, int(p.seed), 3, frames, None, None, None, None, c)
        assert c.cpu().tolist() == want.tolist(), dtype
