"""Bounds discipline of the C ABI without a sanitizer (compute-sanitizer is closed on this GPU
pool): every input sits between NaN guard regions, every output and the workspace between
sentinel regions; after the calls the results must be finite/correct (no guard was read into a
result) and every sentinel untouched (nothing written outside the declared extents)."""
import numpy as np
import pytest
import torch

from helpers import mixed_err
from oracle import closed_form as cf

pytestmark = pytest.mark.gpu
GUARD = 257            # floats on each side (odd on purpose: breaks 16-byte alignment)
SENT = 12345.678


def guarded(a, fill):
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
    buf = torch.full((a.size + 2 * GUARD,), fill, dtype=torch.float32, device="cuda")
    buf[GUARD:GUARD + a.size] = torch.from_numpy(a).cuda()
    return buf, buf[GUARD:GUARD + a.size]


def out_buf(n):
    buf = torch.full((n + 2 * GUARD,), SENT, dtype=torch.float32, device="cuda")
    return buf, buf[GUARD:GUARD + n]


def untouched(buf, n):
    return bool((buf[:GUARD] == SENT).all()) and bool((buf[GUARD + n:] == SENT).all())


@pytest.mark.parametrize("kind,K,N,Mf,Mx", [("gaussian", 1, 2051, 333, 20), ("gaussian", 1, 1, 3, 0),
                                           ("softplus", 1, 515, 70, 9), ("gaussian", 3, 777, 129, 0)])
@pytest.mark.parametrize("sparse", [False, True])
def test_guard_regions(kind, K, N, Mf, Mx, sparse):
    if sparse and (kind != "gaussian" or K != 1):
        pytest.skip("compact-support kernels: gaussian, K=1")
    from spinn_b200 import _cabi
    lib = _cabi.load()
    rs = np.random.RandomState(N)
    M = Mf + Mx
    h0 = 1.0 / np.sqrt(M)
    c = dict(x=rs.rand(N), y=rs.rand(N), xf=rs.rand(Mf), yf=rs.rand(Mf), xfix=rs.rand(Mx), yfix=rs.rand(Mx),
             h=h0 * (1 + 0.3 * rs.rand(M)), W=0.3 * rs.randn(K, M), b=0.1 * rs.randn(K), g=rs.randn(5, N, K))
    nan = float("nan")
    bufs = {k: guarded(v, nan) for k, v in c.items()}
    P = lambda k: bufs[k][1].data_ptr() if bufs[k][1].numel() else None
    kid = cf.kind_id(kind)
    s = torch.cuda.current_stream().cuda_stream
    jets_buf, jets = out_buf(5 * N * K)
    if sparse:
        nb = lib.spinn2d_sparse_fwd_workspace_bytes(N, M)
    else:
        nb = lib.spinn2d_fwd_workspace_bytes(N, M, K)
    ws_buf = torch.full((nb + 2 * 1024,), 0x5A, dtype=torch.uint8, device="cuda")
    ws = ws_buf[1024:1024 + nb]
    assert ws.data_ptr() % 16 == 0
    if sparse:
        rc = lib.spinn2d_jets_fwd_sparse(kid, 2, N, Mf, Mx, K, P("x"), P("y"), P("xf"), P("yf"), P("xfix"), P("yfix"),
                                         P("h"), 0, P("W"), P("b"), None, 7.5, jets.data_ptr(), None,
                                         ws.data_ptr(), nb, s)
    else:
        rc = lib.spinn2d_jets_fwd(kid, 2, N, Mf, Mx, K, P("x"), P("y"), P("xf"), P("yf"), P("xfix"), P("yfix"),
                                  P("h"), 0, P("W"), P("b"), jets.data_ptr(), ws.data_ptr(), nb, s)
    torch.cuda.synchronize()
    assert rc == 0, lib.spinn_last_error()
    assert untouched(jets_buf, 5 * N * K)
    assert bool((ws_buf[:1024] == 0x5A).all()) and bool((ws_buf[1024 + nb:] == 0x5A).all())
    xc, yc = np.concatenate([c["xf"], c["xfix"]]), np.concatenate([c["yf"], c["yfix"]])
    ref = cf.jets2d(kind, c["x"], c["y"], xc, yc, c["h"], c["W"], c["b"])[:5]
    got = jets.cpu().numpy().reshape(5, N, K)
    assert np.all(np.isfinite(got))
    for a in range(5):
        assert mixed_err(got[a], ref[a]) < 5e-5

    outs = {k: out_buf(n) for k, n in (("d_xc", Mf), ("d_yc", Mf), ("d_h", M), ("d_W", K * M), ("d_b", K))}
    O = lambda k: outs[k][1].data_ptr() if outs[k][1].numel() else None
    nb = lib.spinn2d_sparse_bwd_workspace_bytes(N, M) if sparse else lib.spinn2d_bwd_workspace_bytes(N, M, K)
    ws_buf = torch.full((nb + 2 * 1024,), 0x5A, dtype=torch.uint8, device="cuda")
    ws = ws_buf[1024:1024 + nb]
    if sparse:
        rc = lib.spinn2d_jets_bwd_sparse(kid, 2, 0, N, Mf, Mx, K, P("x"), P("y"), P("xf"), P("yf"), P("xfix"),
                                         P("yfix"), P("h"), 0, P("W"), P("g"), None, 7.5, O("d_xc"), O("d_yc"),
                                         O("d_h"), O("d_W"), O("d_b"), None, ws.data_ptr(), nb, s)
    else:
        rc = lib.spinn2d_jets_bwd(kid, 2, 0, N, Mf, Mx, K, P("x"), P("y"), P("xf"), P("yf"), P("xfix"), P("yfix"),
                                  P("h"), 0, P("W"), P("g"), O("d_xc"), O("d_yc"), O("d_h"), O("d_W"), O("d_b"),
                                  ws.data_ptr(), nb, s)
    torch.cuda.synchronize()
    assert rc == 0, lib.spinn_last_error()
    assert bool((ws_buf[:1024] == 0x5A).all()) and bool((ws_buf[1024 + nb:] == 0x5A).all())
    G = cf.grads2d(kind, c["x"], c["y"], xc, yc, c["h"], c["W"], c["g"], n_free=Mf)
    for k, n in (("d_xc", Mf), ("d_yc", Mf), ("d_h", M), ("d_W", K * M), ("d_b", K)):
        assert untouched(outs[k][0], n), k
        v = outs[k][1].cpu().numpy()
        assert np.all(np.isfinite(v)), k
        assert mixed_err(v, np.asarray(G[k]).reshape(-1)) < 6e-5, (k, mixed_err(v, np.asarray(G[k]).reshape(-1)))
