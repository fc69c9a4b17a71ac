"""TEST INFRASTRUCTURE: analytic expectation of the electron images.

Given the integer counts an exposure drew, the per-pixel EXPECTED number of
electrons follows from the reference's recipe alone (pyparallel_menu.c:87-108):
the first (int)(counts*ratio) electrons of a bin are N(x_pos, sigma_h) x
N(y_pos, sigma_h), the others N(., sigma_l), and an electron lands in pixel
(int)x, (int)y.  So

    E[pixel (r, c)] = sum_bins  n_h * Px_h[c] * Py_h[r]  +  n_l * Px_l[c] * Py_l[r]
    P?_x[c] = Phi((c + 1 - x_pos) / sigma) - Phi((c - x_pos) / sigma)

times the sub-sample's flat value at the pixel (grism.py:359-385).  Evaluated in
float64 with torch (erf + batched matrix products) -- on the GPU when there is
one, because the benchmark shape has 1.7e7 (sub-sample, bin) cells -- and shares
no code with the kernels it checks.
"""
import math

import numpy as np
import torch

from oracle import exposure_oracle as E


def _cdf_diff(lo, n, centre, sigma):
    """[B][W][n]: probability of the n unit cells starting at integer `lo`."""
    edges = torch.arange(lo, lo + n + 1, dtype=torch.float64, device=centre.device)
    z = (edges[None, None, :] - centre[:, :, None]) / (sigma[None, :, None] * math.sqrt(2.0))
    c = torch.special.erf(z)
    return 0.5 * (c[:, :, 1:] - c[:, :, :-1])


def flat_value(cal, grism_name, subarray, x_ref, y_ref, rows, cols):
    """Flat value the reference multiplies hit pixel (row, col) by, for a sub-sample
    whose reference position is (x_ref, y_ref): oracle.flat_field_at_hits's
    expression on an explicit pixel list (float32 storage included)."""
    a, b = (E.G141_TRACE, E.G141_WLSOL) if grism_name == 'G141' else (E.G102_TRACE, E.G102_WLSOL)
    f0, f1, f2, f3 = (np.asarray(p) for p in cal['flat'])
    n = len(f0)
    off = (1014 - subarray) // 2
    Y, X = rows + off, cols + off
    Y = np.where(Y < 0, Y + n, Y)
    X = np.where(X < 0, X + n, X)
    m_t, _, m_w, c_w = E.wavelength_calibration_coeffs(x_ref, y_ref, a, b)
    a_t_i = 1 / m_t
    arr = y_ref - Y + a_t_i * x_ref - a_t_i * X
    d = np.sqrt((arr * arr) / (a_t_i * a_t_i + 1))
    w = ((m_w * d + c_w) - cal['flat_wmin']) / (cal['flat_wmax'] - cal['flat_wmin'])
    val = f0[Y, X] + (f1[Y, X] * w) + (f2[Y, X] * (w * w)) + (f3[Y, X] * (w * w * w))
    return val.astype(np.float32).astype(np.float64)


def expected_interval_images(counts, x_pos, y_pos, ratio, sigl, sigh, read_index, L, cal=None,
                             grism_name='G141', subarray=1024, refs=None, batch=16, device=None,
                             zmax=8.5, with_var=False):
    """E[R][L][L] (float64 numpy): expected electrons per read interval and pixel.
    with_var: also the exact variance V[R][L][L] -- the electrons of a bin are multinomial over
    the pixels, so Var = sum n p (1 - p) (flat^2-weighted), a few per cent below E in the
    trace's core where one pixel takes a sizeable share of a bin.

    counts [N][W] int; x_pos, y_pos [N][W] frame coordinates of the bins (already
    minus sub_scale); ratio, sigl, sigh [W]; read_index = last sub-sample of each
    read.  With ``cal`` the sub-sample's flat (refs = [N][2] reference positions)
    multiplies its image."""
    dev = torch.device(device or ('cuda' if torch.cuda.is_available() else 'cpu'))
    N, W = counts.shape
    R = len(read_index)
    out = torch.zeros((R, L, L), dtype=torch.float64, device=dev)
    var = torch.zeros((R, L, L), dtype=torch.float64, device=dev) if with_var else None
    cnt = torch.as_tensor(np.ascontiguousarray(counts), device=dev).to(torch.float64)
    nh = torch.floor(cnt * torch.as_tensor(ratio, device=dev)[None, :])      # (int)(counts*ratio), counts >= 0
    nh = torch.minimum(torch.clamp(nh, min=0.0), cnt)
    nl = cnt - nh
    xs = torch.as_tensor(np.ascontiguousarray(x_pos), device=dev)
    ys = torch.as_tensor(np.ascontiguousarray(y_pos), device=dev)
    sl = torch.as_tensor(np.ascontiguousarray(sigl), device=dev)
    sh = torch.as_tensor(np.ascontiguousarray(sigh), device=dev)
    if cal is not None:
        flat_planes = [torch.as_tensor(np.asarray(p, dtype=np.float64), device=dev) for p in cal['flat']]
        a, b = (E.G141_TRACE, E.G141_WLSOL) if grism_name == 'G141' else (E.G102_TRACE, E.G102_WLSOL)
        nflat = flat_planes[0].shape[0]
        off = (1014 - subarray) // 2
    first = 0
    for r, last in enumerate(read_index):
        for s0 in range(first, last + 1, batch):
            s1 = min(s0 + batch, last + 1)
            sl_b = slice(s0, s1)
            for n_e, sig in ((nh[sl_b], sh), (nl[sl_b], sl)):
                if float(n_e.sum()) == 0.0:
                    continue
                live = n_e > 0
                m = float(sig.max()) * zmax
                x_lo = int(math.floor(float(xs[sl_b][live].min()) - m))
                x_hi = int(math.floor(float(xs[sl_b][live].max()) + m)) + 1
                y_lo = int(math.floor(float(ys[sl_b][live].min()) - m))
                y_hi = int(math.floor(float(ys[sl_b][live].max()) + m)) + 1
                px = _cdf_diff(x_lo, x_hi - x_lo, xs[sl_b], sig)             # [B][W][NX]
                py = _cdf_diff(y_lo, y_hi - y_lo, ys[sl_b], sig)             # [B][W][NY]
                img = torch.bmm((py * n_e[:, :, None]).transpose(1, 2), px)  # [B][NY][NX]
                img2 = torch.bmm((py * py * n_e[:, :, None]).transpose(1, 2), px * px) if with_var else None
                # frame test 0 < x < nr, 0 < y < nc (pyparallel_menu.c:93): clip the box
                cx0, cx1 = max(x_lo, 1), min(x_hi, L)
                cy0, cy1 = max(y_lo, 1), min(y_hi, L)
                if cx0 >= cx1 or cy0 >= cy1:
                    continue
                img = img[:, cy0 - y_lo:cy1 - y_lo, cx0 - x_lo:cx1 - x_lo]
                if with_var:
                    img2 = img2[:, cy0 - y_lo:cy1 - y_lo, cx0 - x_lo:cx1 - x_lo]
                if cal is not None:
                    rows = torch.arange(cy0, cy1, device=dev)
                    cols = torch.arange(cx0, cx1, device=dev)
                    Y = torch.remainder(rows + off, nflat)
                    X = torch.remainder(cols + off, nflat)
                    ref = torch.as_tensor(np.asarray(refs[s0:s1], dtype=np.float64), device=dev)
                    xr, yr = ref[:, 0], ref[:, 1]
                    m_t = (a[3] + a[4] * xr + a[5] * yr + a[6] * xr ** 2 + a[7] * xr * yr + a[8] * yr ** 2)
                    m_w = (b[3] + b[4] * xr + b[5] * yr + b[6] * xr ** 2 + b[7] * xr * yr + b[8] * yr ** 2)
                    c_w = (b[0] + b[1] * xr) + b[2] * yr
                    ati = 1 / m_t
                    arr = (yr[:, None, None] - Y[None, :, None].to(torch.float64)
                           + ati[:, None, None] * (xr[:, None, None] - X[None, None, :].to(torch.float64)))
                    d = torch.sqrt(arr * arr / (ati * ati + 1)[:, None, None])
                    w = ((m_w[:, None, None] * d + c_w[:, None, None]) - cal['flat_wmin']) / (
                        cal['flat_wmax'] - cal['flat_wmin'])
                    f = [p[Y][:, X][None] for p in flat_planes]
                    val = f[0] + f[1] * w + f[2] * w * w + f[3] * w * w * w
                    val = val.to(torch.float32).to(torch.float64)
                    if with_var:                       # Var(flat * n) = flat^2 (n p - n p^2)
                        var[r, cy0:cy1, cx0:cx1] += ((img - img2) * val * val).sum(dim=0)
                    img = img * val
                elif with_var:
                    var[r, cy0:cy1, cx0:cx1] += (img - img2).sum(dim=0)
                out[r, cy0:cy1, cx0:cx1] += img.sum(dim=0)
        first = last + 1
    if with_var:
        return out.cpu().numpy(), var.cpu().numpy()
    return out.cpu().numpy()
