"""Host-side mirror of edu.umass.nlp.optimize.LBFGSMinimizer as Run_NTN uses it
(/root/reference/src/Run_NTN.java:119-124): ``res = LBFGSMinimizer().minimize(t, theta, opts)``,
``res.minArg / res.minObjVal / res.didConverge``.  The iterations run on the device
(csrc/lbfgs.cu): theta is uploaded once, every trial point is one ntn_eval_dev, the two-loop vector
ops are device kernels, and only scalars come back until the final download."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Opts:
    maxIters: int = 5          # LBFGSMinimizer.Opts.maxIters (Run_NTN.java:121 sets batch_iterations = 5)
    maxHistorySize: int = 6
    tol: float = 1e-4


@dataclass
class Result:
    minArg: np.ndarray
    minObjVal: float
    didConverge: bool
    iters: int = 0
    evals: int = 0
    firstObjVal: float = 0.0


class LBFGSMinimizer:
    Opts = Opts

    def minimize(self, fn, theta, opts: Opts | None = None, reduce=None) -> Result:
        """fn: the NTN function object (neural_tensor_network_for_kb_completion_b200.ntn.NTN).  Data parallel: ``reduce``
        (distributed.make_lbfgs_reduce) all-reduces every evaluation's gradient + {cost, n_active} tail, so the replicated
        optimisers of all ranks take identical steps."""
        import torch
        opts = opts or Opts()
        fn._sync_batch()
        h = fn.handle
        # theta stays in HBM between calls: the device copy is reused when the caller passes back the very array the
        # previous call returned (Run_NTN.java:124 does exactly that: theta = res.minArg)
        cache = getattr(fn, "_lbfgs_dev", None)
        if cache is None or cache[0].numel() != h.Pint:
            cache = [torch.empty(h.Pint, dtype=torch.float32, device=f"cuda:{fn.device}"), None]
            fn._lbfgs_dev = cache
        d_theta = cache[0]
        if cache[1] is None or cache[1] is not theta:
            h.upload_theta(np.asarray(theta, np.float64), d_theta.data_ptr())
        # corrupt sides: the seeded Math.random() stream is advanced natively (no interpreter call per evaluation) when
        # the function object offers its state; any other object is called back per evaluation
        st = fn.nativeSidesState() if hasattr(fn, "nativeSidesState") else None
        r = h.lbfgs_minimize(d_theta.data_ptr(), st if st is not None else fn.drawCorruptSides, opts.maxIters,
                             opts.maxHistorySize, opts.tol, reduce=reduce)
        if st is not None:
            fn.adoptSidesState(st)
        fn.update += int(r.n_active_last)
        out = h.download_theta(d_theta.data_ptr())
        out.setflags(write=False)          # identity-keyed reuse of the device copy: an in-place edit would go unnoticed
        cache[1] = out
        return Result(out, r.min_obj_val, bool(r.did_converge), r.iters, r.evals, r.first_cost)
