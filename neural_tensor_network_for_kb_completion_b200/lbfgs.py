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

    def minimize(self, fn, theta, opts: Opts | None = None) -> Result:
        """fn: the NTN function object (neural_tensor_network_for_kb_completion_b200.ntn.NTN)."""
        import torch
        opts = opts or Opts()
        fn._sync_batch()
        h = fn.handle
        d_theta = torch.empty(h.Pint, dtype=torch.float32, device=f"cuda:{fn.device}")
        h.upload_theta(np.asarray(theta, np.float64), d_theta.data_ptr())
        r = h.lbfgs_minimize(d_theta.data_ptr(), fn.drawCorruptSides, opts.maxIters, opts.maxHistorySize, opts.tol)
        fn.update += int(r.n_active_last)
        return Result(h.download_theta(d_theta.data_ptr()), r.min_obj_val, bool(r.did_converge), r.iters, r.evals,
                      r.first_cost)
