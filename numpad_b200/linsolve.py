"""The Newton linear solve on the device (seam S3).

Replaces `scipy.sparse.linalg.spsolve(J, rhs, use_umfpack=False)` of the
reference's Newton loop (adsolve.py:193-195) and the `factorized(...)` solves
of the implicit tangent / adjoint (adsolve.py:78-82, 141-145) with restarted
flexible GMRES (csrc/gmres.cu) on the CUDA SpMV, preconditioned according to
the structure of the matrix.  Everything stays in HBM; there is no CPU path.

Accuracy contract: the TRUE residual is driven to `rtol * ||b||`
(default 1e-12, SuperLU-class) so that Newton trajectories match the
reference's to ~1e-10 (hard part 2).
"""
import ctypes

import torch

from . import _lib
from ._lib import KrylovOps, CsrView, KrylovInfo
from ._dev import DevCsr, device, stream_ptr, _empty

# knobs (also settable through solve(..., linear_options={...}))
DEFAULTS = dict(rtol=1e-12, atol=0.0, restart=100, maxiter=600, precond='auto',
                verbose=False,
                # Krylov method of the preconditioned solves: 'fgmres' (flexible GMRES, csrc/gmres.cu)
                # or 'bicgstab' (csrc/bicgstab.cu; not for the saddle-point preconditioners, whose
                # inner CG makes them non-linear); precond 'block-jacobi' takes block=(bs, layout)
                krylov='fgmres', block=(4, 'strided'),
                # multi-GPU: owner[i] = rank of unknown i.  When set and torch.distributed runs
                # with more than one rank, saddle-point systems too large for the direct solver
                # are solved row-sharded over the ranks (shard.py); every rank calls solve()
                # with the same replicated A and b and gets the same x back.
                shard_owner=None)

# solve_multi: from this many right-hand sides on, the block-tridiagonal direct solver is
# preferred over per-column Krylov solves
MULTI_RHS_DIRECT = 8

last_info = {}      # filled by every solve: iterations, residuals, method
# 'robust': the multigrid-preconditioned solve failed on a matrix of signature 'sig' = (n, nnz):
# later systems of the same signature (the next Newton iterations and time steps of the same
# run) go straight to the robust solver instead of failing first
_hints = {'robust': False, 'sig': None}
_stale = {'btlu': None, 'sig': None, 'transposed': False}     # last block-tridiagonal factorisation (reused as a
                                         # preconditioner while the Jacobian has not moved far)


def reset_hints(hard=True):
    """forget that earlier Jacobians needed the robust solver.  The Newton loop calls this
    softly (hard=False) at the start of every solve: the hint then survives as long as the
    systems keep their signature."""
    if hard:
        _hints['robust'], _hints['sig'] = False, None
        _stale['btlu'], _stale['sig'] = None, None


def forget_factors():
    """drop the block-tridiagonal factorisation kept for reuse as a preconditioner (the
    robust-solver hint itself stays): the next robust solve factorises again"""
    _stale['btlu'], _stale['sig'] = None, None


class LinearSolveError(RuntimeError):
    pass


def _csr_view(A):
    v = CsrView()
    v.nrows, v.nnz = A.shape[0], A.nnz
    v.indptr, v.indices, v.data = A.indptr.data_ptr(), A.indices.data_ptr(), A.data.data_ptr()
    return v


def fgmres(A, b, x0=None, precond=None, rtol=1e-12, atol=0.0, restart=60, maxiter=3000,
           flexible=True, allreduce=None, matvec=None, matvec_c=None, allreduce_c=None):
    """Solve A x = b on the device.  `A` is a DevCsr (or None when `matvec` is
    given); `precond`, `matvec` are python callables (x_ptr, y_ptr) wrapped as
    C callbacks, or ints holding the address of a C callback + ctx pair."""
    L = _lib.lib()
    n = int(b.numel())
    restart = max(1, min(int(restart), n))
    x = torch.zeros(n, dtype=torch.float64, device=device()) if x0 is None else x0.clone()
    ops = KrylovOps()
    keep = []
    if matvec_c is not None:              # (address of an npb_apply_fn, address of its ctx)
        ops.matvec, ops.matvec_ctx = matvec_c
    elif matvec is None:
        A = A.materialized()
        view = _csr_view(A)
        keep += [A, view]
        ops.matvec = L.csr_matvec_cb_addr
        ops.matvec_ctx = ctypes.addressof(view)
    else:
        cb = _lib.APPLY_FN(_wrap_apply(matvec, n))
        keep.append(cb)
        ops.matvec = ctypes.cast(cb, ctypes.c_void_p).value
    if precond is not None and hasattr(precond, 'c_apply'):
        keep.append(precond)             # C callback + ctx: no Python in the loop
        ops.precond, ops.precond_ctx = precond.c_apply, precond.c_ctx
    elif precond is not None:
        cb = _lib.APPLY_FN(_wrap_apply(precond, n))
        keep.append(cb)
        ops.precond = ctypes.cast(cb, ctypes.c_void_p).value
    if allreduce_c is not None:           # (address of an npb_allreduce_fn, its ctx)
        ops.allreduce, ops.allreduce_ctx = allreduce_c
    elif allreduce is not None:
        cb = _lib.ALLREDUCE_FN(allreduce)
        keep.append(cb)
        ops.allreduce = ctypes.cast(cb, ctypes.c_void_p).value
    flexible = 1 if (flexible and precond is not None) else 0
    wbytes = L.cdll.npb_fgmres_workspace_bytes(n, restart, flexible)
    work = _empty(wbytes, torch.uint8)
    info = KrylovInfo()
    rc = L.cdll.npb_fgmres(n, ctypes.addressof(ops), b.data_ptr(), x.data_ptr(), restart,
                           int(maxiter), float(rtol), float(atol), flexible,
                           work.data_ptr(), wbytes, ctypes.addressof(info), stream_ptr())
    res = dict(status=rc, iters=info.iters, restarts=info.restarts, bnorm=info.bnorm,
               rnorm0=info.rnorm0, rnorm=info.rnorm)
    if rc not in (0, -5, -4):
        raise RuntimeError('npb_fgmres failed (%d): %s' % (rc, L.last_error()))
    return x, res


def bicgstab(A, b, x0=None, precond=None, rtol=1e-12, atol=0.0, maxiter=3000, allreduce_c=None,
             matvec_c=None):
    """Solve A x = b on the device with right-preconditioned BiCGStab (csrc/bicgstab.cu); same
    conventions as fgmres().  The preconditioner must be a fixed linear operator."""
    L = _lib.lib()
    n = int(b.numel())
    x = torch.zeros(n, dtype=torch.float64, device=device()) if x0 is None else x0.clone()
    ops = KrylovOps()
    keep = []
    if matvec_c is not None:
        ops.matvec, ops.matvec_ctx = matvec_c
    else:
        A = A.materialized()
        view = _csr_view(A)
        keep += [A, view]
        ops.matvec = L.csr_matvec_cb_addr
        ops.matvec_ctx = ctypes.addressof(view)
    if precond is not None and hasattr(precond, 'c_apply'):
        keep.append(precond)
        ops.precond, ops.precond_ctx = precond.c_apply, precond.c_ctx
    elif precond is not None:
        cb = _lib.APPLY_FN(_wrap_apply(precond, n))
        keep.append(cb)
        ops.precond = ctypes.cast(cb, ctypes.c_void_p).value
    if allreduce_c is not None:
        ops.allreduce, ops.allreduce_ctx = allreduce_c
    wbytes = L.cdll.npb_bicgstab_workspace_bytes(n)
    work = _empty(wbytes, torch.uint8)
    info = KrylovInfo()
    rc = L.cdll.npb_bicgstab(n, ctypes.addressof(ops), b.data_ptr(), x.data_ptr(), int(maxiter),
                             float(rtol), float(atol), work.data_ptr(), wbytes,
                             ctypes.addressof(info), stream_ptr())
    res = dict(status=rc, iters=info.iters, restarts=info.restarts, bnorm=info.bnorm,
               rnorm0=info.rnorm0, rnorm=info.rnorm)
    if rc not in (0, -5, -4):
        raise RuntimeError('npb_bicgstab failed (%d): %s' % (rc, L.last_error()))
    return x, res


_cb_error = []


def _wrap_apply(fn, n):
    """python callable (x: tensor, y: tensor) -> C callback on raw pointers"""
    def cb(ctx, xptr, yptr, stream):
        try:
            x = _tensor_from_ptr(xptr, n)
            y = _tensor_from_ptr(yptr, n)
            fn(x, y)
            return 0
        except Exception as e:            # never unwind through C
            _cb_error.append(e)
            return -2
    return cb


def _tensor_from_ptr(ptr, n):
    """zero-copy float64 view of n doubles at device address `ptr`"""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {'shape': (n,), 'typestr': '<f8', 'data': (int(ptr), False),
                                  'version': 3, 'strides': None}
    return torch.as_tensor(h, device=device())


def _now():
    import time
    torch.cuda.synchronize()
    return time.perf_counter()


def _solve_sharded(A, b, o):
    """the row-sharded path of solve(); None when it does not apply or did not converge (every
    rank takes the same decision: all of them see the same matrix and the same iteration)"""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return None
    from . import precond as _pc, amg, shard
    if _pc.direct_feasible(A) or not bool((amg.diagonal(A) == 0).any()):
        return None
    try:
        x = shard.solve(A, b, owner=o['shard_owner'], rtol=o['rtol'], restart=max(o['restart'], 100),
                        maxiter=2 * max(o['restart'], 100))
    except LinearSolveError:
        _hints['robust'] = True        # convection-dominated: the replicated robust solver takes over
        return None
    last_info.clear()
    last_info.update(shard.last_info)
    if o['verbose']:
        print('    linear solve [%s]: %d its, |r|/|b| = %.2e' %
              (last_info['method'], last_info['iters'],
               last_info['rnorm'] / max(last_info['bnorm'], 1e-300)))
    return x


def _btlu_refined(M, A, b, o, transposed=False):
    """x from the block-tridiagonal factors M of A (of A^T when `transposed`: then M's
    transposed sweeps are used) + FGMRES refinement"""
    pc = M.apply_t if transposed else M.apply
    x0 = M.solve(b, transpose=transposed)
    # refinement: a few preconditioned steps towards 1e-2 of the requested tolerance (a direct
    # solve is expected to deliver SuperLU-class residuals, and each step is one pair of sweeps);
    # accepted at the requested tolerance
    x, res = fgmres(A, b, x0=x0, precond=pc, rtol=max(1e-2 * o['rtol'], 2e-15), atol=o['atol'],
                    restart=8, maxiter=8)
    if res['rnorm'] <= max(o['rtol'] * res['bnorm'], o['atol']):
        res['status'] = 0
    return x, res


def _solve_btlu(A, b, o, t0):
    """block-tridiagonal direct solve + FGMRES refinement to the requested tolerance; a
    factorisation of an earlier Jacobian of the same pattern is first tried as the
    preconditioner (Newton iterates move little: a handful of iterations instead of a new
    factorisation).  None: not applicable (caller falls back)."""
    from . import blocktri, precond as _pc
    import numpy as np
    sig = (A.shape[0], A.nnz)
    info = {}
    M = _stale['btlu'] if _stale['sig'] == sig else None
    if M is not None and o.get('btlu_reuse', True):
        # (for A = J^T the factors of J serve through their transposed sweeps)
        use_t = bool(o.get('_transposed')) != bool(_stale['transposed'])
        # each iteration is one pair of sweeps over the stored inverses (~25 ms at 1024^2 against
        # ~15 s for a new factorisation), so the stale factors get a generous budget
        x, res = fgmres(A, b, precond=M.apply_t if use_t else M.apply, rtol=o['rtol'],
                        atol=o['atol'], restart=60, maxiter=o.get('btlu_stale_iters', 120))
        if res['status'] == 0:
            if o.get('_keep') is not None:
                o['_keep'].update(M=M, method='btlu-stale', A=A, use_t=use_t)
            res.update(method='btlu (reused factorisation)', setup_ms=0.0, krylov_ms=1e3 * (_now() - t0),
                       precond=M.describe())
            last_info.clear()
            last_info.update(res)
            if o['verbose']:
                print('    linear solve [%s]: %d its, |r|/|b| = %.2e' % (res['method'], res['iters'],
                      res['rnorm'] / max(res['bnorm'], 1e-300)))
            return x
        info['stale_iters'] = res['iters']
    _stale['btlu'] = M = None            # free the old inverses before building new ones
    if not blocktri.feasible(A):
        return None
    try:
        M = blocktri.BlockTriLU(A)
    except np.linalg.LinAlgError:
        return None
    t_setup = _now() - t0
    t1 = _now()
    x, res = _btlu_refined(M, A, b, o)
    res.update(info, method='btlu', setup_ms=1e3 * t_setup, krylov_ms=1e3 * (_now() - t1),
               precond=M.describe())
    last_info.clear()
    last_info.update(res)
    if _cb_error:
        e = _cb_error.pop()
        _cb_error.clear()
        raise e
    if o['verbose']:
        print('    linear solve [btlu]: %sfactor %.0f ms, %d refinement its, |r|/|b| = %.2e' %
              ('stale factors gave up after %d its, ' % info['stale_iters'] if 'stale_iters' in info else '',
               res['setup_ms'], res['iters'], res['rnorm'] / max(res['bnorm'], 1e-300)))
    if res['status'] != 0:
        return None
    _stale['btlu'], _stale['sig'], _stale['transposed'] = M, sig, bool(o.get('_transposed'))
    if o.get('_keep') is not None:
        o['_keep'].update(M=M, method='btlu', A=A)
    return x


def solve(A, b, transpose=False, **opts):
    """x = A^{-1} b (or A^{-T} b) for a DevCsr A and a device vector b."""
    o = dict(DEFAULTS)
    o.update(opts)
    keep = o.get('_keep')
    b = b.reshape(-1).contiguous()
    if keep is not None and 'M' in keep:
        return _solve_again(keep, b, o)
    if transpose:
        A = A.transpose()
        o['_transposed'] = True
    A = A.materialized()
    n = A.shape[0]
    from . import precond as _pc
    if o['shard_owner'] is not None and not _hints['robust'] and o['precond'] in ('auto', 'lsc'):
        x = _solve_sharded(A, b, o)
        if x is not None:
            return x
    t0 = _now()
    sig = (n, A.nnz)
    if _hints['robust'] and _hints['sig'] != sig:
        reset_hints()
    if _stale['btlu'] is not None and _stale['sig'] != sig:
        forget_factors()          # a different system: do not sit on tens of GB of old inverses
    if o['precond'] in ('auto', 'lsc', 'amg') and _hints['robust'] and not _pc.direct_feasible(A):
        o['precond'] = _pc.robust_kind(A)
    if o['precond'] == 'btlu':
        x = _solve_btlu(A, b, o, t0)
        if x is not None:
            return x
        o['precond'] = 'lsc-robust' if bool((_pc.diagonal(A) == 0).any()) else 'jacobi'
    M, method = _pc.build(A, o['precond'], block=o['block'])
    first_try = method in ('fgmres+lsc', 'fgmres+amg')
    t_setup = _now() - t0
    if method == 'direct':
        x = M.solve(b)
        r = b - A.matvec(x)
        res = dict(status=0, iters=0, rnorm=float(r.norm()), bnorm=float(b.norm()))
        # one step of iterative refinement keeps the residual at SuperLU level
        if res['rnorm'] > o['rtol'] * max(res['bnorm'], 1e-300):
            x = x + M.solve(r)
            res['rnorm'] = float((b - A.matvec(x)).norm())
        res['method'] = method
        res['setup_ms'] = 1e3 * t_setup
        last_info.clear()
        last_info.update(res)
        if keep is not None:
            keep.update(M=M, method=method, A=A)
        return x
    t0 = _now()
    pc = None if M is None else (M if hasattr(M, 'c_apply') else M.apply)
    # a multigrid-preconditioned solve that has not converged within two restart
    # cycles will not: fail fast and let the robust retry below take over
    maxit = min(o['maxiter'], 2 * o['restart']) if first_try else \
        (max(o['maxiter'], 2000) if method == 'fgmres+lsc-robust' else o['maxiter'])
    if o['krylov'] == 'bicgstab' and not method.startswith('fgmres+lsc'):
        x, res = bicgstab(A, b, precond=pc, rtol=o['rtol'], atol=o['atol'], maxiter=maxit)
        method = method.replace('fgmres', 'bicgstab').replace('gmres', 'bicgstab')
    else:
        x, res = fgmres(A, b, precond=pc, rtol=o['rtol'],
                        atol=o['atol'], restart=min(o['restart'], n), maxiter=maxit)
    res['method'] = method
    res['setup_ms'] = 1e3 * t_setup
    res['krylov_ms'] = 1e3 * (_now() - t0)
    if hasattr(M, 'describe'):
        res['precond'] = M.describe()
    last_info.clear()
    last_info.update(res)
    if _cb_error:
        e = _cb_error.pop()
        _cb_error.clear()
        raise e
    if o['verbose']:
        print('    linear solve [%s]: %d its, |r|/|b| = %.2e' %
              (method, res['iters'], res['rnorm'] / max(res['bnorm'], 1e-300)))
    if res['status'] != 0 and method in ('fgmres+lsc', 'fgmres+amg'):
        # multigrid on a convection-dominated momentum block / a non-elliptic (compressible-
        # flow) Jacobian can stagnate or diverge: retry with the robust solver -- the block-
        # tridiagonal direct solve when its dense blocks fit in HBM, else the Krylov inner
        # solve (saddle points) or Jacobi-GMRES -- and keep using it for the systems that follow
        _hints['robust'], _hints['sig'] = True, sig
        del M
        kind = _pc.robust_kind(A)
        return solve(A, b, **dict(o, precond=kind,
                                  maxiter=max(o['maxiter'], 1000) if kind == 'jacobi' else o['maxiter']))
    if res['status'] != 0 and method != 'direct' and _pc.direct_feasible(A):
        # robust fallback: pivoted banded LU (+ refinement through the front door)
        return solve(A, b, **dict(o, precond='direct'))
    if res['status'] != 0:
        raise LinearSolveError('linear solve (%s) did not converge: %d iterations, '
                               '|r|/|b| = %.3e' % (method, res['iters'],
                                                  res['rnorm'] / max(res['bnorm'], 1e-300)))
    if keep is not None:
        keep.update(M=M, method=method, A=A, maxit=maxit)
    return x


def _solve_again(keep, b, o):
    """another right-hand side for the matrix of an earlier solve(..., _keep=keep): the built
    factorisation / preconditioner is applied again, nothing is set up"""
    M, method, A = keep['M'], keep['method'], keep['A']
    t0 = _now()
    if method == 'direct':
        x = M.solve(b)
        r = b - A.matvec(x)
        res = dict(status=0, iters=0, rnorm=float(r.norm()), bnorm=float(b.norm()))
        if res['rnorm'] > o['rtol'] * max(res['bnorm'], 1e-300):
            x = x + M.solve(r)
            res['rnorm'] = float((b - A.matvec(x)).norm())
    elif method == 'btlu':
        x, res = _btlu_refined(M, A, b, o)
    elif method == 'btlu-stale':
        x, res = fgmres(A, b, precond=M.apply_t if keep.get('use_t') else M.apply, rtol=o['rtol'],
                        atol=o['atol'], restart=30, maxiter=60)
    else:
        pc = None if M is None else (M if hasattr(M, 'c_apply') else M.apply)
        x, res = fgmres(A, b, precond=pc, rtol=o['rtol'], atol=o['atol'],
                        restart=min(o['restart'], A.shape[0]), maxiter=keep.get('maxit', o['maxiter']))
    res.update(method=method + ' (solver reused)', setup_ms=0.0, krylov_ms=1e3 * (_now() - t0))
    last_info.clear()
    last_info.update(res)
    if res['status'] != 0:
        raise LinearSolveError('linear solve (%s, reused) did not converge: |r|/|b| = %.3e'
                               % (method, res['rnorm'] / max(res['bnorm'], 1e-300)))
    return x


def solve_multi(A, B, transpose=False, **opts):
    """X with A X[k] = B[k] (or A^T X[k] = B[k]) for the rows of the dense device matrix B
    (k x n): the reference loops one LU solve per row / column of a densified operand
    (adsolve.py:74-82, 137-145); here the solver -- factorisation or preconditioner hierarchy
    -- is built ONCE for the first right-hand side and reused for the others."""
    keep = {}
    X = torch.empty_like(B)
    if B.shape[0] >= MULTI_RHS_DIRECT and opts.get('precond', DEFAULTS['precond']) == 'auto':
        # many right-hand sides: one factorisation + one pair of sweeps per column beats one
        # preconditioned Krylov solve per column (when the dense blocks fit)
        from . import blocktri, precond as _pc
        Am = A.materialized()
        if not _pc.direct_feasible(Am) and blocktri.feasible(Am):
            opts = dict(opts, precond='btlu')
    for k in range(B.shape[0]):
        X[k] = solve(A, B[k], transpose=transpose, **dict(opts, _keep=keep))
    return X
