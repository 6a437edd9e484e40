"""One training step of the image INR fit, natively fused.

Mirrors `ImageTrainer.step` (wisp/trainers/image_trainer.py:269-359) for the static-coordinate image fit:

    feats = grid.interpolate(coords)            latent_grid.py:340-382   -> tiled forward kernel
    rgb   = decoder_color(feats)                nefs/image.py:109-120,152 \\  fused MLP + MSE kernel
    rgb_loss = ((rgb - gt) ** 2).mean()         image_trainer.py:298-300  /  (value + every gradient)
    ent   = grid.ent_loss(it)[0]                latent_grid.py:122-136   -> fused bit-rate kernel
    loss  = rgb_loss + lambda * ent             image_trainer.py:314-319
    loss.backward(); optimizer.step()           :321-359, parameter groups base_trainer.py:206-266

In PyTorch terms that step is ~50 kernels (autograd glue, gradient scaling passes over the table, one multi-tensor
Adam per group). Here it is TWO launches (+ one 2.4 KB memset) and no autograd:
  * shacira_fit_tile_step -- grid forward + decoder MLP + MSE + grid backward in one tile-resident kernel
    (csrc/fit_kernels.cuh): the [N, 16] feature rows and their gradient never leave the SM;
  * shacira_fit_optimizer_step -- everything table-side in one launch (csrc/optimizer_kernels.cuh): the bit-rate loss of
    the latents (value + gradients; the latents' bit-rate gradient goes straight into the table's Adam, lambda is a device
    scalar), Adam over the table and over all ~20 small tensors with their chain rules, and the NEXT step's SGA sample of
    the freshly updated latents.
Shapes outside the fused kernels' (latent_dim / feature_dim != 1, levels that do not fit a tile, injected SGA draws) keep
the separate launches: tiled forward -> tensor-core MLP -> tiled backward, bit-rate kernel on a forked stream, SGA kernel.
Grid and MLP exchange rows in the plan's tile order (sorted-I/O plan, targets permuted once), and with device_noise=True
the bit-rate noise is drawn in the kernel. The whole step is CUDA-graph capturable; `set_lambda` / `set_temperature` /
`set_sga` / `draw_noise` / `update_div` are the host-side schedule hooks of the trainer (image_trainer.py:131-137,284-296).

The module parameters stay the single source of truth: the kernels read and update the storage of
`grid.codebook`, `grid.latent_dec.layers[0].{scale,shift}`, `grid.prob_model.f*.{h,b,a}` and the MLP in place, so
`grid.interpolate`, `grid.size()`, `state_dict()` etc. see the trained values at any time.
"""
import ctypes
import os

import torch

from . import _lib, grid_ops


class ImageFitStep:
    def __init__(self, grid, mlp, coords, target, lr=1e-3, grid_lr=2e-2, ldec_lr=1e-2, prob_lr=1e-4,
                 weight_decay=0.0, weight_decay_decoder=1e-2, betas=(0.9, 0.999), eps=1e-8, device_noise=False,
                 noise_seed=0):
        """`grid`: shacira_b200.grids.LatentGrid (2D, single affine decoder; SGA sampling or STE rounding as the decoder's
        `use_sga` says -- see set_sga / set_temperature); `mlp`:
        nn.Sequential(Linear(L*F,16), ReLU, Linear(16,16), ReLU, Linear(16,3)); coords [N,2], target [N,3].
        Learning rates / weight decays: the reference's parameter groups (base_trainer.py:219-239; kodak.yaml:61-70).
        device_noise=True draws the bit-rate noise inside the kernel (fresh on every step, also under graph replay:
        `draw_noise` is then unnecessary); False reads `self.noise`, which `draw_noise` fills (parity runs)."""
        dec = grid.latent_dec
        amap = dec.affine_map() if hasattr(dec, "affine_map") else None
        if amap is None or amap[0].shape[0] != 1 or "dft" in dec.layers[0].ldecode_matrix:
            raise _lib.ShaciraError(_lib.ERR_UNSUPPORTED, "ImageFitStep: needs ONE affine 'sq' latent decoder")
        if grid.prob_model is None:
            raise _lib.ShaciraError(_lib.ERR_UNSUPPORTED, "ImageFitStep: needs the bit-rate model (entropy_reg > 0)")
        self.grid, self.mlp = grid, mlp
        self.coords = _lib._f32c(coords, "coords")
        self.target = _lib._f32c(target, "target")
        dev = self.coords.device
        self.dev = dev
        self.n = self.coords.shape[0]
        self.T, self.C = grid.codebook.shape
        self.L = grid.num_lods
        layer = dec.layers[0]
        self.F = layer.scale.shape[1]
        self.fi, _ = _lib._i32_array(grid._first_idx())
        self.rs, _ = _lib._i32_array(grid.resolutions)
        self.bw = int(grid.codebook_bitwidth)
        self.betas, self.eps = betas, float(eps)
        self.grid_lr, self.weight_decay = float(grid_lr), float(weight_decay)
        self.lin = [mlp[0], mlp[2], mlp[4]]
        self.IN, self.H, self.OUT = self.lin[0].in_features, self.lin[0].out_features, self.lin[2].out_features
        if self.IN != self.L * self.F:
            raise _lib.ShaciraError(_lib.ERR_INVALID_ARGUMENT, "MLP input width must be num_lods * feature_dim")
        if self.n < 4096 or self.L % 4:
            raise _lib.ShaciraError(_lib.ERR_UNSUPPORTED, "ImageFitStep: coordinate set / level count outside the tiled path")
        # A plan of its own, in sorted-I/O mode: the step's consumers of the feature rows (per-point MLP, mean loss)
        # do not care about the order of the points, so grid and MLP exchange rows in the plan's tile order (targets
        # permuted once here) and the grid kernels lose their perm -> row dependent loads.
        self.plan = _lib.Plan(self.coords).set_sorted_io(True)
        self.target = self.target[self.plan.perm_tensor()].contiguous()
        f32 = dict(dtype=torch.float32, device=dev)
        # density-model parameters live in ONE [4, 3, C] buffer (the kernel's layout); the module's h / b / a become
        # views of it, so nothing is packed per step
        pm = grid.prob_model
        self.prob = torch.zeros((4, 3, self.C), **f32)
        for i, f in enumerate((pm.f1, pm.f2, pm.f3, pm.f4)):
            for k, name in enumerate(("h", "b", "a")):
                p = getattr(f, name, None)
                if p is None:
                    continue
                self.prob[i, k].copy_(p.detach().reshape(-1))
                p.data = self.prob[i, k].view(1, self.C)
        self.num_prob_layers = int(pm.num_layers)
        # step buffers
        self.A = torch.empty((1, self.C, self.F), **f32)
        self.feats = torch.empty((self.n, self.L * self.F), **f32)
        self.gfeat = torch.empty_like(self.feats)
        self.use_bound = self.IN == 16 and os.environ.get("SHACIRA_MLP_IMPL", "tc") == "tc"   # the tensor-core kernel
        # one tile-resident kernel for grid forward + MLP / MSE + grid backward where its shapes apply (SHACIRA_FIT_FUSED=0:
        # the three-kernel path, kept for every other shape and as the parity reference of the fused kernel)
        self.opt_fused = os.environ.get("SHACIRA_FIT_OPT_FUSED", "1") != "0"   # one optimizer launch (incl. next SGA sample)
        self._what_valid = False
        # the bit-rate loss evaluated inside the optimizer launch's table pass (latent_dim 1): no bit-rate launch, no
        # forked stream, the latents' bit-rate gradient never goes through memory
        self.ent_in_opt = self.opt_fused and self.C == 1 and os.environ.get("SHACIRA_FIT_ENT_FUSED", "1") != "0"
        self.fused = (os.environ.get("SHACIRA_FIT_FUSED", "1") != "0" and self.use_bound and self.H == 16 and
                      self.OUT == 3 and self.C == 1 and self.F == 1 and self.L == 16 and amap[0].shape[0] == 1)
        n_par = self.H * self.IN + self.H + self.H * self.H + self.H + self.OUT * self.H + self.OUT
        # double SSE | packed MLP gradients | max |feature gradient| per column (reduced by the MLP kernel; kept behind
        # the gradients so that the call clears everything with one memset)
        self.mlp_out = torch.zeros(2 + n_par + 16, **f32)
        self.gfeat_max = self.mlp_out[2 + n_par:]
        # accumulated into by the backward, cleared by the table's Adam kernel after use: no memset per step
        self.g_grid = torch.zeros((self.T, self.C), **f32)
        self.g_ent = torch.empty((self.T, self.C), **f32)
        self.g_prob = torch.zeros((4, 3, self.C), **f32)
        self.g_dec = torch.zeros((self.L * self.C * self.F + self.L * self.F,), **f32)   # dA rows | dshift rows
        self.bits = torch.zeros((1 + self.L,), dtype=torch.float64, device=dev)
        self.noise = torch.zeros((self.T, self.C), **f32)
        self.lam = torch.zeros((), **f32)
        self.device_noise, self.noise_seed = bool(device_noise), int(noise_seed)
        self.rng_step = torch.zeros((), dtype=torch.int64, device=dev)
        # SGA (basic_latent_decoder.py:183-191): the reference's quantiser until epoch / max_epochs > decay_period
        # (image_trainer.py:136-137). One table-side kernel per step writes w_hat (what the grid kernels interpolate,
        # rounding off) and d w_hat / d w (multiplied into the grid gradient inside the table's Adam kernel).
        self.sga = bool(getattr(dec, "use_sga", False))
        self.diff_sampling = bool(getattr(dec, "diff_sampling", True))
        self.temperature = torch.full((), float(getattr(dec, "temperature", 1.0)), **f32)
        self.w_hat = torch.empty((self.T, self.C), **f32)
        self.dw = torch.empty((self.T, self.C), **f32)
        self.sga_uniforms = None     # [T, C, 2]: injected draws (parity runs); None = drawn in the kernel
        self.sga_rng_step = torch.zeros((), dtype=torch.int64, device=dev)
        self.ent_scratch = torch.zeros(int(_lib.load().shacira_entropy_scratch_bytes(self.C, self.L)),
                                       dtype=torch.uint8, device=dev)
        # Adam state
        self.m_table, self.v_table = torch.zeros_like(grid.codebook.data), torch.zeros_like(grid.codebook.data)
        self.step_table = torch.zeros((), **f32)
        self.step_small = torch.zeros((), **f32)
        self.adam_ticket = torch.zeros((), dtype=torch.int32, device=dev)   # arrival counter of the per-tensor CTAs
        self._keep = []
        # the bit-rate kernel depends on nothing but the table: it runs on a forked stream beside the grid / MLP
        # kernels (it is latency bound: 17 us alone) and joins before the optimizer kernels
        self.side = torch.cuda.Stream(device=dev)
        self._forked = torch.cuda.Event()
        segs = []

        def seg(param, grad, n, lr, wd, rows=1, stride=0, scale=None, mul=1.0, div=None, group=1, zero=False):
            m, v = torch.zeros(n, **f32), torch.zeros(n, **f32)
            self._keep += [m, v]
            s = _lib.AdamSeg()
            s.param, s.grad = param.data_ptr(), grad.data_ptr()
            s.exp_avg, s.exp_avg_sq = m.data_ptr(), v.data_ptr()
            s.grad_scale = scale.data_ptr() if scale is not None else None
            s.grad_div = div.data_ptr() if div is not None else None
            s.n, s.grad_rows, s.grad_row_stride, s.div_group = n, rows, stride, group
            s.lr, s.weight_decay, s.grad_mul, s.zero_grad = lr, wd, mul, 1.0 if zero else 0.0
            segs.append(s)

        packed = self.mlp_out[2:2 + n_par]
        off = 0
        for lin in self.lin:                       # decoder group: lr, weight decay 0 (base_trainer.py:222-224)
            for p in (lin.weight, lin.bias):
                seg(p.data, packed[off:], p.numel(), float(lr), 0.0)
                off += p.numel()
        CF = self.C * self.F
        # latent_dec group (base_trainer.py:225-229): scale = A * div  =>  dscale = (sum_l dA_l) / div
        seg(layer.scale.data, self.g_dec, CF, float(ldec_lr), float(weight_decay_decoder), rows=self.L, stride=CF,
            div=dec.div.data, group=self.F, zero=True)
        self.has_shift = layer.shift is not None
        if self.has_shift:
            seg(layer.shift.data, self.g_dec[self.L * CF:], self.F, float(ldec_lr), float(weight_decay_decoder),
                rows=self.L, stride=self.F, zero=True)
        # prob_model group: lr fixed 1e-4 (base_trainer.py:230-234); gradient = lambda / rows * d(bits)
        used = [0] if self.num_prob_layers > 1 else []
        if self.num_prob_layers > 2:
            used.append(1)
        if self.num_prob_layers > 3:
            used.append(2)
        used.append(3)
        for i, f in zip(range(4), (pm.f1, pm.f2, pm.f3, pm.f4)):
            if i not in used:
                continue                            # never receives a gradient: torch.optim.Adam skips it too
            for k, name in enumerate(("h", "b", "a")):
                p = getattr(f, name, None)
                if p is None or not p.requires_grad:
                    continue
                seg(self.prob[i, k], self.g_prob[i, k], self.C, float(prob_lr), float(weight_decay_decoder),
                    scale=self.lam, mul=1.0 / self.T)
        if len(segs) > _lib.MAX_ADAM_SEGS:
            raise _lib.ShaciraError(_lib.ERR_UNSUPPORTED, "too many small tensors")
        self.segs = (_lib.AdamSeg * len(segs))(*segs)
        self.nseg = len(segs)
        self.layer = layer
        self.refresh_A()

    # ---- host-side schedule hooks ------------------------------------------------------------
    def refresh_A(self):
        """A = scale / div (latent_decoders.affine_map); call after changing scale or div outside step()."""
        with torch.no_grad():
            self.A.copy_((self.layer.scale.data / self.grid.latent_dec.div.data.unsqueeze(1)).unsqueeze(0))

    def set_lambda(self, value):
        self.lam.fill_(float(value))

    def set_temperature(self, value, refresh=False):
        """SGA temperature of this epoch (image_trainer.py:131-133; base_trainer.py:155-157 builds the schedule). The
        sample of the NEXT step has already been drawn by the previous optimizer launch with the temperature of that
        moment: refresh=True discards it, so the new value applies from the very next step (an epoch-boundary caller);
        without it the new value applies one step later (a per-step schedule under CUDA-graph replay)."""
        self.temperature.fill_(float(value))
        self.grid.latent_dec.temperature = float(value)
        if refresh:
            self._what_valid = False

    def set_sga(self, flag):
        """Switch between SGA and straight-through rounding (the trainer turns SGA off once epoch / max_epochs >
        decay_period). A captured CUDA graph holds the mode it was captured with: re-capture after switching."""
        self.sga = bool(flag)
        self.grid.latent_dec.use_sga = bool(flag)
        self._what_valid = False

    def draw_noise(self, generator=None):
        """U(-0.5, 0.5) per latent (latent_grid.py:128); with a CPU generator the reference's stream."""
        if generator is not None:
            self.noise.copy_(torch.rand((self.T, self.C), generator=generator) - 0.5)
        else:
            self.noise.uniform_(-0.5, 0.5)

    def update_div(self):
        """norm='max' (image_trainer.py:284-296): div = max(|min|, |max|) per latent channel."""
        with torch.no_grad():
            w = self.grid.codebook.data
            self.grid.latent_dec.div.data.copy_(torch.max(torch.abs(w.min(dim=0)[0]), torch.abs(w.max(dim=0)[0])))
        self.refresh_A()

    # ---- the step --------------------------------------------------------------------------------
    def step(self):
        lib, P, chk = _lib.load(), _lib._ptr, _lib._check
        g, dec = self.grid, self.grid.latent_dec
        lat = g.codebook.data
        shift = self.layer.shift.data if self.has_shift else None
        lin = self.lin
        with torch.cuda.device(self.dev):
            st = _lib._stream()
            cur = torch.cuda.current_stream(self.dev)
            self._forked.record(cur)   # the table as the previous optimizer launch left it
            q_lat, rflag, gmul = lat, 1, None
            # SGA sample of this step: normally already there -- the previous step's optimizer launch sampled the latents it
            # had just updated (shacira_fit_optimizer_step); drawn here on the first step, after set_sga / a refreshing
            # set_temperature, and whenever the draws are injected (parity runs)
            prefetch = self.sga and self.opt_fused and self.sga_uniforms is None
            if self.sga:
                if not (prefetch and self._what_valid):
                    chk(lib.shacira_sga_quantize(P(lat), P(self.sga_uniforms), self.T * self.C, P(self.temperature),
                                                 1 if self.diff_sampling else 0, self.noise_seed + 0x5A17,
                                                 P(self.sga_rng_step), P(self.w_hat), P(self.dw), st))
                q_lat, rflag, gmul = self.w_hat, 0, self.dw
            CF = self.C * self.F
            if not self.has_shift:
                self.g_dec[self.L * CF:].zero_()   # no segment consumes (and clears) the shift rows
            fused_done = False
            if self.fused:
                # grid forward + MLP / MSE + grid backward as ONE tile-resident kernel (csrc/fit_kernels.cuh): the feature
                # rows and their gradient never leave the SM
                rc = lib.shacira_fit_tile_step(self.plan.handle, P(q_lat), self.fi, self.rs, self.L, self.bw, rflag,
                                               P(self.A), P(shift), P(self.target), P(lin[0].weight.data),
                                               P(lin[0].bias.data), P(lin[1].weight.data), P(lin[1].bias.data),
                                               P(lin[2].weight.data), P(lin[2].bias.data), self.T, P(self.g_grid),
                                               P(self.g_dec), P(self.g_dec[self.L * CF:]), P(self.mlp_out), st)
                if rc == _lib.ERR_UNSUPPORTED:
                    self.fused = False     # e.g. a level whose node box does not fit a tile: the three-kernel path
                else:
                    chk(rc)
                    fused_done = True
            if not fused_done:
                chk(lib.shacira_latent_forward_planned(self.plan.handle, P(q_lat), self.fi, self.rs, self.L, self.bw, self.C,
                                                       self.F, rflag, P(self.A), P(shift), 0, P(self.feats), st))
                # the MLP kernel reduces max |feature gradient| per column on the way; the tiled backward takes its
                # fixed-point scales from that bound and skips its own pass over the gradient rows
                bound = P(self.gfeat_max) if self.use_bound else None
                chk(lib.shacira_mlp_mse_step_bounded(P(self.feats), P(self.target), self.n, self.IN, self.H, self.OUT,
                                                     P(lin[0].weight.data), P(lin[0].bias.data), P(lin[1].weight.data),
                                                     P(lin[1].bias.data), P(lin[2].weight.data), P(lin[2].bias.data),
                                                     P(self.gfeat), None, P(self.mlp_out), bound, st))
                chk(lib.shacira_latent_backward_planned_bounded(self.plan.handle, P(self.gfeat), P(q_lat), self.fi, self.rs,
                                                                self.L, self.bw, self.C, self.F, rflag, P(self.A), 0, self.T, 0,
                                                                P(self.g_grid), P(self.g_dec), P(self.g_dec[self.L * CF:]),
                                                                bound, st))
            # The bit-rate loss: normally evaluated inside the optimizer launch's table pass (ent_in_opt). Otherwise its own
            # kernel on a forked stream, launched AFTER the tile kernel (it only depends on the table).
            ent_folded = self.ent_in_opt and self.opt_fused
            if not ent_folded:
                self.side.wait_event(self._forked)
                with torch.cuda.stream(self.side):
                    if self.device_noise:
                        chk(lib.shacira_entropy_bits_rng(P(lat), self.noise_seed, P(self.rng_step), self.T, self.C,
                                                         P(self.prob), self.num_prob_layers, self.fi, self.L, P(self.bits),
                                                         P(self.g_ent), P(self.g_prob), P(self.ent_scratch),
                                                         self.ent_scratch.numel(), _lib._stream()))
                    else:
                        chk(lib.shacira_entropy_bits(P(lat), P(self.noise), self.T, self.C, P(self.prob),
                                                     self.num_prob_layers, self.fi, self.L, P(self.bits), P(self.g_ent),
                                                     P(self.g_prob), P(self.ent_scratch), self.ent_scratch.numel(),
                                                     _lib._stream()))
                cur.wait_stream(self.side)
            if self.opt_fused:
                # ONE optimizer launch: small tensors + table Adam + the next step's SGA sample (csrc/optimizer_kernels.cuh)
                chk(lib.shacira_fit_optimizer_step(ctypes.cast(self.segs, ctypes.c_void_p), self.nseg, P(lat), P(self.g_grid),
                                                   P(gmul), None if ent_folded else P(self.g_ent), P(self.lam),
                                                   1.0 / self.T, P(self.m_table),
                                                   P(self.v_table), self.T * self.C, self.grid_lr, self.weight_decay,
                                                   self.betas[0], self.betas[1], self.eps, P(self.step_small),
                                                   P(self.step_table), P(self.layer.scale.data), P(dec.div.data), P(self.A),
                                                   self.C, self.F, P(self.temperature), 1 if self.diff_sampling else 0,
                                                   self.noise_seed + 0x5A17, P(self.sga_rng_step),
                                                   P(self.w_hat) if prefetch else None, P(self.dw) if prefetch else None,
                                                   P(self.prob) if ent_folded else None, self.num_prob_layers,
                                                   None if self.device_noise else P(self.noise), self.noise_seed,
                                                   P(self.rng_step), P(self.bits), P(self.g_prob), P(self.ent_scratch),
                                                   self.ent_scratch.numel(), P(self.adam_ticket), st))
                self._what_valid = prefetch
            else:
                chk(lib.shacira_adam_step_sum_mul(P(lat), P(self.g_grid), P(gmul), P(self.g_ent), P(self.lam), 1.0 / self.T,
                                                  P(self.m_table), P(self.v_table), self.T * self.C, self.grid_lr,
                                                  self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                                  P(self.step_table), 0, 1, st))
                chk(lib.shacira_multi_adam_step(ctypes.cast(self.segs, ctypes.c_void_p), self.nseg, self.betas[0],
                                                self.betas[1], self.eps, P(self.step_small), P(self.step_table),
                                                P(self.layer.scale.data), P(dec.div.data), P(self.A), self.C, self.F,
                                                P(self.adam_ticket), st))
                self._what_valid = False

    # ---- results of the last step (device tensors; reading them synchronises) -------------------
    def rgb_loss(self):
        return (self.mlp_out[:2].view(torch.float64)[0] / (self.n * self.OUT)).to(torch.float32)

    def total_bits(self):
        return self.bits[0].to(torch.float32)

    def close(self):
        if self.plan is not None:
            self.plan.close()
            self.plan = None
