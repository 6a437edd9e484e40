"""TEST INFRASTRUCTURE ONLY -- float64 restatement of the packed ray integration the reference's tracer calls
(wisp/tracers/packed_rf_tracer.py:136-153 -> kaolin.render.spc.exponential_integration / sum_reduce, kaolin 0.13.0).

kaolin is absent from this image and from /root/reference (requirements: README.md:35), so this restates the
PUBLISHED algorithm -- per packed ray: T_i = exp(-exclusive cumsum of tau), w_i = T_i (1 - exp(-tau_i)),
ray_feats = sum_i w_i feats_i -- and is anchored on the reference's call site. PARITY WITH KAOLIN ITSELF IS UNPINNED.
Plain loops / torch float64; only tests/, smoke() and bench.py's cpu_baseline leg may import this."""
import numpy as np
import torch


def exponential_integration(feats, tau, boundary):
    """numpy float64: (ray_feats [R, NF], weights [S])."""
    feats = np.asarray(feats, dtype=np.float64)
    tau = np.asarray(tau, dtype=np.float64).reshape(-1)
    starts = list(np.nonzero(np.asarray(boundary).reshape(-1))[0]) + [tau.shape[0]]
    w = np.zeros_like(tau)
    out = np.zeros((len(starts) - 1, feats.shape[1]))
    for r in range(len(starts) - 1):
        acc = 0.0
        for i in range(starts[r], starts[r + 1]):
            w[i] = np.exp(-acc) * (1.0 - np.exp(-tau[i]))
            acc += tau[i]
            out[r] += w[i] * feats[i]
    return out, w


def exponential_integration_torch(feats, tau, boundary):
    """Differentiable float64 torch form of the same thing (autograd gives the reference gradients)."""
    tau = tau.reshape(-1)
    starts = torch.nonzero(boundary.reshape(-1)).squeeze(1).tolist() + [tau.shape[0]]
    outs, ws = [], []
    for r in range(len(starts) - 1):
        t = tau[starts[r]:starts[r + 1]]
        excl = torch.cumsum(t, 0) - t
        w = torch.exp(-excl) * (1.0 - torch.exp(-t))
        ws.append(w)
        outs.append((w.unsqueeze(1) * feats[starts[r]:starts[r + 1]]).sum(0))
    return torch.stack(outs), torch.cat(ws)
