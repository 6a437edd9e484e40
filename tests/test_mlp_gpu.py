"""Fused decoder MLP + MSE kernel against the plain PyTorch fp32 reference of the same op."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,in_dim,ref_dtype", [(1, 16, torch.float32), (31, 16, torch.float32),
                                                 (4096 + 17, 16, torch.float32), (100000, 16, torch.float32),
                                                 (5000, 24, torch.float32), (3001, 32, torch.float32),
                                                 (393216, 16, torch.float64)])
def test_fused_mlp_mse_matches_torch(lib, n, in_dim, ref_dtype):
    """Reference: the same op in plain PyTorch, fp32. At the full Kodak size the reference is evaluated in fp64:
    measured on B200, torch's fp32 Linear at N = 393 216 flips the ReLU sign of pre-activations that are ~1e-7 from
    zero relative to fp64 (13 % max error on single feature-gradient rows), while this kernel agrees with fp64 to
    2e-7 -- so fp64 is the trustworthy yardstick there."""
    from shacira_b200 import grid_ops
    torch.manual_seed(n + in_dim)
    mlp = nn.Sequential(nn.Linear(in_dim, 16), nn.ReLU(), nn.Linear(16, 16), nn.ReLU(), nn.Linear(16, 3)).cuda()
    x = (torch.randn(n, in_dim, device="cuda") * 0.7).requires_grad_(True)
    gt = torch.rand(n, 3, device="cuda")
    if ref_dtype == torch.float64:
        import copy
        ref_mlp = copy.deepcopy(mlp).double()
        xr = x.detach().double().requires_grad_(True)
        pred_ref = ref_mlp(xr)
        loss_ref = ((pred_ref - gt.double()) ** 2).mean()
        (3.0 * loss_ref).backward()
        want = dict(x=xr.grad.clone(), **{k: p.grad.clone() for k, p in ref_mlp.named_parameters()})
        loss, pred = grid_ops.mlp_mse_loss(x, gt, mlp, want_pred=True)
        (3.0 * loss).backward()
        assert abs(loss.item() - loss_ref.item()) <= 1e-5 * abs(loss_ref.item())
        assert rel_err(pred.cpu().numpy(), pred_ref.detach().cpu().numpy()) <= 1e-5
        # Feature gradients go through two ReLU masks: a pre-activation within rounding distance of zero flips its
        # mask and changes that ONE row by O(1) -- in any fp32-grade implementation (torch's own fp32 Linear shows 13 %
        # on such rows, see above). Hold every row to 1e-4 except a handful of mask-flip rows (<= 1 in 10 000), and
        # hold the typical row to 1e-5; the weight gradients (sums over all rows) are held to 1e-4 as everywhere.
        err = (x.grad.double() - want["x"]).abs().amax(dim=1) / want["x"].abs().max()
        assert int((err > 1e-4).sum()) <= n // 10000, int((err > 1e-4).sum())
        assert float(err.quantile(0.999)) <= 1e-5
        for k, p in mlp.named_parameters():
            assert rel_err(p.grad.cpu().numpy(), want[k].cpu().numpy()) <= 1e-4, k
        return
    pred_ref = mlp(x)
    loss_ref = ((pred_ref - gt) ** 2).mean()
    (3.0 * loss_ref).backward()
    want = dict(x=x.grad.clone(), **{k: p.grad.clone() for k, p in mlp.named_parameters()})
    x.grad = None
    mlp.zero_grad()
    loss, pred = grid_ops.mlp_mse_loss(x, gt, mlp, want_pred=True)
    (3.0 * loss).backward()
    assert abs(loss.item() - loss_ref.item()) <= 1e-5 * abs(loss_ref.item())
    assert rel_err(pred.cpu().numpy(), pred_ref.detach().cpu().numpy()) <= 1e-5
    assert rel_err(x.grad.cpu().numpy(), want["x"].cpu().numpy()) <= 1e-4
    for k, p in mlp.named_parameters():
        assert rel_err(p.grad.cpu().numpy(), want[k].cpu().numpy()) <= 1e-4, k


def test_fused_mlp_rejects_other_shapes(lib):
    from shacira_b200 import grid_ops
    mlp = nn.Sequential(nn.Linear(16, 32), nn.ReLU(), nn.Linear(32, 32), nn.ReLU(), nn.Linear(32, 3)).cuda()
    with pytest.raises(lib.ShaciraError):
        grid_ops.mlp_mse_loss(torch.zeros(8, 16, device="cuda"), torch.zeros(8, 3, device="cuda"), mlp)


def test_table_adam_matches_torch_adam(lib):
    from shacira_b200._lib import TableAdam
    torch.manual_seed(0)
    for n, wd in ((374612, 0.0), (1001, 0.01)):
        p_ref = torch.nn.Parameter(torch.randn(n, 1, device="cuda"))
        p_our = torch.nn.Parameter(p_ref.detach().clone())
        ref = torch.optim.Adam([p_ref], lr=2e-2, weight_decay=wd)
        our = TableAdam(p_our, lr=2e-2, weight_decay=wd)
        for it in range(5):
            g = torch.randn(n, 1, device="cuda") * (10.0 ** (it - 2))
            p_ref.grad = g.clone()
            p_our.grad = g.clone()
            ref.step()
            our.step()
        assert rel_err(p_our.detach().cpu().numpy(), p_ref.detach().cpu().numpy()) <= 1e-6
        assert float(our.step_count) == 5.0
