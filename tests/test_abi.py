"""CPU-side checks of the boundary: the C-ABI library builds, loads, and exports exactly the symbols
include/b200phy.h declares; the ctypes structs match the C layout; the product refuses to run
without a GPU instead of falling back to the CPU oracle."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'b200phy.h')


def declared_symbols():
    hdr = open(HEADER).read()
    return sorted(set(re.findall(r'\b(b200phy_[a-z0-9_]+)\s*\(', hdr)))


def test_library_exports_every_declared_symbol():
    from pyphysim_b200 import _build, _lib
    _build.build()
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.SIGNATURES) == names          # binding covers the whole header, nothing else
    assert lib.b200phy_version() == 101
    assert lib.b200phy_launch_count() == 0           # no compute without a GPU


def test_param_struct_layout_matches_c(tmp_path):
    from pyphysim_b200 import _lib
    src = tmp_path / 'sz.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "b200phy.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu\\n", sizeof(b200phy_ofdm_tdl_params),'
                   'offsetof(b200phy_ofdm_tdl_params, delays), offsetof(b200phy_ofdm_tdl_params, tap_powers),'
                   'offsetof(b200phy_ofdm_tdl_params, seed), sizeof(b200phy_modem));return 0;}\n')
    exe = tmp_path / 'sz'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
    size, o_del, o_pow, o_seed, msize = (int(v) for v in subprocess.check_output([str(exe)]).split())
    P = _lib.OfdmTdlParams
    assert (C.sizeof(P), P.delays.offset, P.tap_powers.offset, P.seed.offset) == (size, o_del, o_pow, o_seed)
    assert C.sizeof(_lib.Modem) == msize


def test_header_is_plain_c():
    # the boundary must be bindable from C / cgo / JNI: compile the header as C89-ish C
    subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Werror', '-fsyntax-only', '-x', 'c', HEADER])


def test_product_does_not_import_oracle():
    bad = []
    for root, _, files in os.walk(os.path.join(ROOT, 'pyphysim_b200')):
        for f in files:
            if f.endswith('.py'):
                txt = open(os.path.join(root, f)).read()
                if re.search(r'^\s*(from|import)\s+oracle\b', txt, re.M):
                    bad.append(f)
    assert not bad, bad


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from pyphysim_b200 import _lib
    from pyphysim_b200.modulators import fundamental as F
    q = F.QAM(16)
    with pytest.raises(_lib.B200PhyError):
        q.modulate([0, 1, 2])
    with pytest.raises(_lib.B200PhyError):
        from pyphysim_b200 import links
        links.link_siso_flat(q, 0.1, 16)


def test_argument_validation_needs_no_gpu():
    """Bad arguments are rejected by the C layer before any launch (reference error semantics)."""
    from pyphysim_b200 import _lib
    lib = _lib.load()
    m = _lib.Modem(_lib.MODEM_QAM, 32, 1)
    cnt = (C.c_int64 * 4)()
    rc = lib.b200phy_link_siso_flat(_lib.F32, m, 1, 0.1, 1, 0, 16, None, None, None, None, None,
                                    C.cast(cnt, C.c_void_p), None)
    assert rc == _lib.ERR_INVALID and b'square power of 2' in lib.b200phy_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)
    m = _lib.Modem(_lib.MODEM_QAM, 16, 1)
    rc = lib.b200phy_link_alamouti(_lib.F32, m, 2, 3, 0.1, 1, 0, 16, None, None, None, None, None,
                                   C.cast(cnt, C.c_void_p), None)
    assert rc == _lib.ERR_INVALID
    rc = lib.b200phy_link_blast(_lib.F32, m, 2, 2, 1, -1.0, 0.0, 1, 0, 16, None, None, None, None, None,
                                C.cast(cnt, C.c_void_p), None)
    assert rc == _lib.ERR_INVALID and b'non-negative' in lib.b200phy_last_error()
    p = _lib.OfdmTdlParams()
    p.struct_size = C.sizeof(p)
    p.fft, p.cp, p.used, p.n_sym, p.Nr, p.Nt, p.n_taps, p.L = 64, 65, 52, 1, 1, 1, 1, 8
    p.Ts, p.tap_powers[0] = 1e-6, 1.0
    rc = lib.b200phy_link_ofdm_tdl(C.byref(p), m, 0, 1, None, None, None, None, None, None, None,
                                   C.cast(cnt, C.c_void_p), None)
    assert rc == _lib.ERR_INVALID and b'cp_size' in lib.b200phy_last_error()
    p.cp, p.used = 16, 51
    rc = lib.b200phy_link_ofdm_tdl(C.byref(p), m, 0, 1, None, None, None, None, None, None, None,
                                   C.cast(cnt, C.c_void_p), None)
    assert rc == _lib.ERR_INVALID and b'multiple of 2' in lib.b200phy_last_error()
    # the parameter check alone (what the host-buffer entry points run before they size any copy)
    assert lib.b200phy_ofdm_tdl_check_params(C.byref(p)) == _lib.ERR_INVALID
    p.used = 52
    assert lib.b200phy_ofdm_tdl_check_params(C.byref(p)) == 0
    p.struct_size = 8
    assert lib.b200phy_ofdm_tdl_check_params(C.byref(p)) == _lib.ERR_INVALID and b'size mismatch' in lib.b200phy_last_error()
    p.struct_size = C.sizeof(p)
    for nr in range(1, 6):                                # every Nt <= Nr <= 4 is built, nothing else
        for nt in range(1, 6):
            p.Nr, p.Nt = nr, nt
            want = 0 if nt <= nr <= 4 else _lib.ERR_UNSUPPORTED
            assert lib.b200phy_ofdm_tdl_check_params(C.byref(p)) == want, (nr, nt)
    # host entry points validate before touching the device: bad arguments fail here even without a GPU
    cnt64 = (C.c_int64 * 4)()
    p.Nr, p.Nt, p.n_taps = 1, 1, 99
    rc = lib.b200phy_link_ofdm_tdl_host(C.byref(p), _lib.MODEM_QAM, 16, None, 0, 4, None, None, None, None, None, cnt64)
    assert rc in (_lib.ERR_INVALID, _lib.ERR_UNSUPPORTED) and b'n_taps' in lib.b200phy_last_error()
    tab = (C.c_double * 8)()
    rc = lib.b200phy_link_alamouti_host(_lib.F32, _lib.MODEM_QPSK, 4, tab, 2, 3, 0.1, 1, 0, 8, None, None, None, None, cnt64)
    assert rc == _lib.ERR_INVALID and b'even' in lib.b200phy_last_error()
    one = (C.c_uint8 * 8)()
    rc = lib.b200phy_link_blast_host(_lib.F32, _lib.MODEM_QPSK, 4, tab, 2, 2, 1, 0.1, 0.0, 1, 0, 8,
                                     C.cast(one, C.c_void_p), None, None, None, cnt64)
    assert rc == _lib.ERR_INVALID and b'together' in lib.b200phy_last_error()
    rc = lib.b200phy_link_precoded_host(_lib.F32, _lib.MODEM_QPSK, 4, tab, 9, 2, 2, 1, 0.1, 0.0, 1, 0, 8, None, None,
                                        None, None, cnt64)
    assert rc == _lib.ERR_INVALID and b'scheme' in lib.b200phy_last_error()
    rc = lib.b200phy_link_siso_flat_host(_lib.F32, _lib.MODEM_QAM, 16, None, 1, 0.1, 1, 0, 8, None, None, None, None, cnt64)
    assert rc == _lib.ERR_INVALID and b'table_re_im' in lib.b200phy_last_error()
