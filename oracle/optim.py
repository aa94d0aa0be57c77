"""Oracle (test infrastructure, see oracle/__init__.py): numpy restatement of the objective of
CameraGroup.optim_points (reference aniposelib/cameras.py):
  interpolate_data / medfilt_data          :129-145   pre-processing of the start track
  _initialize_params_triangulation         :1664-1690 start values of the limb lengths
  _error_fun_triangulation(_jointlenfix)   :1560-1620, :1356-1416  the residual vector
Pinned against the executed reference by tests/golden/optim_*.npz (oracle/make_golden.py case_optim)."""
import numpy as np

from . import cameragroup as og


def interpolate_data(vals):
    """:138-145 — linear interpolation over NaNs, constant beyond the ends; all-NaN -> 0."""
    nans = np.isnan(vals)
    out = np.copy(vals)
    if nans.all():
        out[:] = 0
    elif nans.any():
        ix = np.arange(vals.size)
        out[nans] = np.interp(ix[nans], ix[~nans], vals[~nans])
    return out


def medfilt_data(values, size=7):
    """:129-133 — reflect padding by size + 5, scipy.signal.medfilt, crop."""
    from scipy import signal
    padsize = size + 5
    vpad = np.pad(values, (padsize, padsize), mode="reflect")
    return signal.medfilt(vpad, kernel_size=size)[padsize:-padsize]


def scale_smooth_full(p3ds, scale_smooth):
    """:1148-1153."""
    intp = np.apply_along_axis(interpolate_data, 0, p3ds)
    med = np.apply_along_axis(medfilt_data, 0, intp, size=7)
    return scale_smooth * (1.0 / np.mean(np.abs(np.diff(med, axis=0)))), intp


def initialize_params(p3ds, constraints, constraints_weak):
    """:1664-1690."""
    jl = np.array([np.median(np.linalg.norm(p3ds[:, a] - p3ds[:, b], axis=1)) for a, b in constraints])
    jw = np.array([np.median(np.linalg.norm(p3ds[:, a] - p3ds[:, b], axis=1)) for a, b in constraints_weak])
    jl, jw = jl.reshape(-1).astype(float), jw.reshape(-1).astype(float)
    both = np.hstack([jl, jw])
    med = np.median(both)
    if med == 0:
        med = 1e-3
    mad = np.median(np.abs(both - med))
    jl[jl == 0] = med
    jw[jw == 0] = med
    jl[jl > med + mad * 5] = med
    jw[jw > med + mad * 5] = med
    return np.hstack([p3ds.ravel(), jl, jw])


def error_fun(cams, params, p2ds, constraints=(), constraints_weak=(), scores=None, scale_smooth=10000,
              scale_length=1, scale_length_weak=0.2, reproj_error_threshold=100, reproj_loss="soft_l1",
              n_deriv_smooth=1, joint_len=None):
    """:1560-1620 (joint_len given: the _jointlenfix form :1356-1416)."""
    n_cams, n_frames, n_joints, _ = p2ds.shape
    n_3d = n_frames * n_joints * 3
    K, Kw = len(constraints), len(constraints_weak)
    p3ds = params[:n_3d].reshape((n_frames, n_joints, 3))
    if joint_len is None:
        jl, jw = params[n_3d:n_3d + K], params[n_3d + K:]
    else:
        jl, jw = joint_len[:K], joint_len[K:]
    p2f = p2ds.reshape((n_cams, -1, 2))
    errors = og.reprojection_error(cams, p3ds.reshape(-1, 3), p2f)
    if scores is not None:
        errors = errors * scores.reshape((n_cams, -1))[:, :, None]
    e = np.abs(errors[~np.isnan(p2f)])
    rp = reproj_error_threshold
    if reproj_loss == "huber":
        bad = e > rp
        e[bad] = rp * (2 * np.sqrt(e[bad] / rp) - 1)
    elif reproj_loss == "soft_l1":
        e = rp * 2 * (np.sqrt(1 + e / rp) - 1)
    smooth = np.diff(p3ds, n=n_deriv_smooth, axis=0).ravel() * scale_smooth

    def lengths(cons, exp, sc):
        out = np.empty((len(cons), n_frames))
        for i, (a, b) in enumerate(cons):
            ln = np.linalg.norm(p3ds[:, a] - p3ds[:, b], axis=1)
            out[i] = 100 * (ln - exp[i]) / exp[i]
        return out.ravel() * sc
    return np.hstack([e, smooth, lengths(constraints, jl, scale_length), lengths(constraints_weak, jw, scale_length_weak)])


def jac_sparsity(p2ds, constraints=(), constraints_weak=(), n_deriv_smooth=1, fix_lengths=False):
    """Sparsity pattern handed to least_squares (:1714-1793; the _jointlenfix form :1272-1352 drops the
    length columns): which parameters each residual depends on."""
    from scipy.sparse import coo_matrix
    n_cams, n_frames, n_joints, _ = p2ds.shape
    K, Kw = len(constraints), len(constraints_weak)
    flat = p2ds.reshape((n_cams, -1, 2))
    good = ~np.isnan(flat)
    pt = np.broadcast_to(np.arange(flat.shape[1])[None, :, None], flat.shape)[good]
    n_r = int(good.sum())
    n_s = (n_frames - n_deriv_smooth) * n_joints * 3
    n_3d = n_frames * n_joints * 3
    n_params = n_3d + (0 if fix_lengths else K + Kw)
    rows, cols = [], []
    ix = np.arange(n_r)
    for k in range(3):
        rows.append(ix)
        cols.append(pt * 3 + k)
    p3 = np.arange(n_frames * n_joints).reshape((n_frames, n_joints))
    fr = np.arange(n_frames - n_deriv_smooth)
    for j in range(n_joints):
        for n in range(n_deriv_smooth + 1):
            pa, pb = p3[fr, j], p3[fr + n, j]
            for k in range(3):
                rows.append(n_r + pa * 3 + k)
                cols.append(pb * 3 + k)
    fr = np.arange(n_frames)
    start = n_r + n_s
    for cix, (a, b) in enumerate(list(constraints) + list(constraints_weak)):
        r = start + cix * n_frames + fr
        if not fix_lengths:
            rows.append(r)
            cols.append(np.full(n_frames, n_3d + cix))
        for p in (p3[fr, a], p3[fr, b]):
            for k in range(3):
                rows.append(r)
                cols.append(p * 3 + k)
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    n_err = n_r + n_s + (K + Kw) * n_frames
    m = coo_matrix((np.ones(rows.size, dtype=np.int16), (rows, cols)), shape=(n_err, n_params)).tocsr()
    m.data[:] = 1
    return m


def optim_points_port(cams, points, p3ds, constraints=(), constraints_weak=(), scale_smooth=4, scale_length=2,
                      scale_length_weak=0.5, reproj_error_threshold=15, reproj_loss="soft_l1", n_deriv_smooth=1,
                      scores=None, joint_len=None):
    """CameraGroup.optim_points / optim_points_jointlenfix as the reference runs them (:1116-1270):
    scipy.optimize.least_squares (trf, 2-point Jacobian on the sparsity pattern, ftol 1e-3; the fixed-length
    form adds max_nfev = 15).  The CPU baseline ("port") of bench.py's optim line."""
    from scipy import optimize
    s_full, intp = scale_smooth_full(p3ds, scale_smooth)
    x0 = initialize_params(intp, constraints, constraints_weak)
    x0[~np.isfinite(x0)] = 0
    fix = joint_len is not None
    if fix:
        x0 = x0[:p3ds.size]
    jac = jac_sparsity(points, constraints, constraints_weak, n_deriv_smooth, fix)
    fun = lambda x: error_fun(cams, x, points, constraints, constraints_weak, scores, s_full, scale_length,
                              scale_length_weak, reproj_error_threshold, reproj_loss, n_deriv_smooth, joint_len)
    extra = {"max_nfev": 15} if fix else {}
    opt = optimize.least_squares(fun, x0=x0, jac_sparsity=jac, loss="linear", ftol=1e-3, **extra)
    new = opt.x[:p3ds.size].reshape(p3ds.shape)
    return new, (joint_len if fix else opt.x[p3ds.size:]), 0.5 * float(opt.fun @ opt.fun)
