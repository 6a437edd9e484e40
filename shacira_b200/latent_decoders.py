"""Latent decoders of SHACIRA's LatentGrid, host-side mirror of
`wisp/models/latent_decoders/` (basic_latent_decoder.py, hierarchical_latent_decoder.py,
multi_latent_decoder.py): same class names, constructor arguments, parameter names
(`div`, `layers.N.scale`, `layers.N.shift`, `decoders.L...`, `alpha`) and semantics, so
reference checkpoints and config dicts load unchanged.

On the hot path these modules are NOT executed table-wide each step the way the reference
does (`latent_dec(codebook)`, latent_grid.py:359): when the decoder is affine
(`affine_map()` is not None) `LatentGrid.interpolate` folds quantisation and decode into the
fused CUDA kernel. `forward()` remains the table-side definition (used for non-affine
configurations, SGA sampling and by the parity tests).
"""
import math

import torch
import torch.nn as nn

epsilon = 1e-6  # basic_latent_decoder.py:13


def get_dft_matrix(conv_dim, channels):
    """DCT-II style basis, basic_latent_decoder.py:14-21: row i, column j."""
    i = torch.arange(conv_dim, dtype=torch.float64).unsqueeze(1) + 0.5
    j = torch.arange(channels, dtype=torch.float64).unsqueeze(0)
    dft = torch.cos(math.pi / channels * i * j) / math.sqrt(channels)
    dft[:, 1:] *= math.sqrt(2)
    return dft.to(torch.float32)


class StraightThrough(torch.autograd.Function):
    """round() forward (half to even), identity backward. basic_latent_decoder.py:28-36."""

    @staticmethod
    def forward(ctx, x):
        return torch.round(x)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output


class StraightThroughFloor(torch.autograd.Function):
    """floor() forward, identity backward. basic_latent_decoder.py:38-46."""

    @staticmethod
    def forward(ctx, x):
        return torch.floor(x)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output


class SineScaled(nn.Module):
    def __init__(self, w0=30.0):
        super().__init__()
        self.w0 = w0

    def forward(self, x):
        return torch.sin(self.w0 * x)


def _activation(name):
    table = {"none": nn.Identity, "sigmoid": nn.Sigmoid, "tanh": nn.Tanh, "relu": nn.ReLU,
             "sine": lambda: SineScaled(30.0)}
    return table[name]()


FORCE_TORCH_SGA = False   # parity runs: the torch definition below also for CUDA tensors (benchmarks/fit_image.py)


def sga_quantize(weight, temperature, diff_sampling, uniforms=None):
    """Stochastic Gumbel annealing between floor and ceil (basic_latent_decoder.py:183-191), active for the first
    `decay_period` of training (every shipped yaml: use_sga True, decay_period 0.9). CUDA tensors take the fused
    table-side kernel (shacira_sga_quantize: value + derivative in one pass, noise drawn in the kernel or injected
    through `uniforms` [T, C, 2]); its output feeds the fused grid kernels with rounding disabled. The torch
    expression below is the definition (host tensors, multi decoder)."""
    if weight.is_cuda and weight.dtype == torch.float32 and not FORCE_TORCH_SGA:
        from . import grid_ops
        return grid_ops.sga_quantize(weight, temperature, diff_sampling, uniforms=uniforms)
    wf = torch.floor(weight) if diff_sampling else StraightThroughFloor.apply(weight)
    wc = wf + 1
    lo, hi = -1 + epsilon, 1 - epsilon
    logit_f = -torch.tanh(torch.clamp(weight - wf, min=lo, max=hi)).unsqueeze(-1) / temperature
    logit_c = -torch.tanh(torch.clamp(wc - weight, min=lo, max=hi)).unsqueeze(-1) / temperature
    logits = torch.cat((logit_f, logit_c), dim=-1)
    if uniforms is None:
        dist = torch.distributions.relaxed_categorical.RelaxedOneHotCategorical(temperature, logits=logits)
        sample = dist.rsample() if diff_sampling else dist.sample()
    else:
        # RelaxedOneHotCategorical.rsample on given draws (torch/distributions/relaxed_categorical.py: normalised logits,
        # clamp_probs, Gumbels, (logits + g) / temperature, log-softmax, exp)
        logits = logits - logits.logsumexp(dim=-1, keepdim=True)
        eps = torch.finfo(uniforms.dtype).eps
        u = uniforms.reshape(logits.shape).clamp(min=eps, max=1 - eps)
        scores = (logits - (-(u.log())).log()) / temperature
        sample = (scores - scores.logsumexp(dim=-1, keepdim=True)).exp()
        if not diff_sampling:
            sample = sample.detach()
    return wf * sample[..., 0] + wc * sample[..., 1]


class DecoderLayer(nn.Module):
    """One linear map latent -> feature. basic_latent_decoder.py:48-95."""

    def __init__(self, in_features, out_features, ldecode_matrix, bias=False):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.ldecode_matrix = ldecode_matrix
        if "dft" in ldecode_matrix:
            self.dft = nn.Parameter(get_dft_matrix(in_features, out_features), requires_grad=False)
            self.scale = nn.Parameter(torch.empty((1, out_features)))
        else:
            self.scale = nn.Parameter(torch.empty((in_features, out_features)))
        if bias:
            self.shift = nn.Parameter(torch.empty(1, out_features))
        else:
            self.register_parameter("shift", None)
        if ldecode_matrix == "dft_fixed":
            self.scale.requires_grad_(False)

    def reset_parameters(self, param=1.0, init_type="normal"):
        if init_type == "normal":
            nn.init.normal_(self.scale, std=param)
        elif init_type == "uniform":
            nn.init.uniform_(self.scale, -param, param)
        elif init_type == "constant":
            nn.init.constant_(self.scale, val=param)
        if self.shift is not None:
            nn.init.zeros_(self.shift)

    def clamp(self, val=0.5):
        with torch.no_grad():
            self.scale.clamp_(-val, val)

    def matrix(self):
        """[in, out] matrix this layer multiplies by."""
        return self.dft * self.scale if "dft" in self.ldecode_matrix else self.scale

    def forward(self, x):
        if "dft" in self.ldecode_matrix:
            out = torch.matmul(x, self.dft) * self.scale
        else:
            out = torch.matmul(x, self.scale)
        return out + self.shift if self.shift is not None else out

    def extra_repr(self):
        return "in_features={}, out_features={}, bias={}".format(self.in_features, self.out_features,
                                                                 self.shift is not None)


class LatentDecoder(nn.Module):
    """basic_latent_decoder.py:97-198."""

    def __init__(self, latent_dim, feature_dim, norm, ldecode_matrix, use_shift, num_layers_dec=0,
                 hidden_dim_dec=0, activation="none", final_activation="none", clamp_weights=0.0, ldec_std=1.0,
                 use_sga=False, diff_sampling=False, **kwargs):
        super().__init__()
        latent_dim = feature_dim if latent_dim == 0 else latent_dim
        self.ldecode_matrix = ldecode_matrix
        self.channels = feature_dim
        self.latent_dim = latent_dim
        self.norm = norm
        self.div = nn.Parameter(torch.ones(latent_dim), requires_grad=False)
        self.num_layers_dec = num_layers_dec
        if num_layers_dec > 0:
            if hidden_dim_dec == 0:
                hidden_dim_dec = feature_dim
            if not isinstance(hidden_dim_dec, (tuple, list)):
                hidden_dim_dec = (hidden_dim_dec,) * num_layers_dec
            self.hidden_dim_dec = tuple(hidden_dim_dec)
        self.use_shift = use_shift
        self.activation_name = activation
        self.final_activation_name = final_activation
        self.act = _activation(activation)
        self.final_activation = _activation(final_activation)
        self.clamp_weights = clamp_weights

        layers, width = [], latent_dim
        for l in range(num_layers_dec):
            hidden = self.hidden_dim_dec[l] or width
            layers += [DecoderLayer(width, hidden, ldecode_matrix, bias=use_shift), self.act]
            width = hidden
        layers.append(DecoderLayer(width, feature_dim, ldecode_matrix, bias=use_shift))
        self.use_sga = use_sga
        self.temperature = 1.0
        self.sga_uniforms = None   # [T, C, 2] U(0,1) draws to use instead of fresh noise (parity tests)
        self.layers = nn.Sequential(*layers)
        self.reset_parameters("normal", ldec_std)
        self.diff_sampling = diff_sampling

    def reset_parameters(self, init_type, param=0.5):
        for layer in self.layers.children():
            if isinstance(layer, DecoderLayer):
                layer.reset_parameters(param, init_type)

    def get_scale(self):
        assert self.num_layers_dec == 0, "Can only get scale for 0 hidden layers decoder!"
        return self.layers[0].scale

    def clamp(self, val=0.2):
        for layer in self.layers.children():
            if isinstance(layer, DecoderLayer):
                layer.clamp(val)

    def size(self, use_torchac=False):
        return sum(p.numel() * torch.finfo(p.dtype).bits for p in self.parameters())

    def scale_norm(self):
        if self.num_layers_dec > 0:
            print("Warning: norm is not implemented for multiple layer decoder>0, returning default value 1")
            return 1
        return self.layers[0].scale.norm()

    def scale_grad_norm(self):
        if self.num_layers_dec > 0:
            print("Warning: norm is not implemented for multiple layer decoder>0, returning default value 1")
            return 1
        return self.layers[0].scale.grad.norm()

    # -- fused-path interface ------------------------------------------------------------
    def is_affine(self):
        return (self.num_layers_dec == 0 and self.final_activation_name == "none"
                and not (self.clamp_weights > 0.0))

    def affine_map(self):
        """(A[1, C, F], shift[1, F] | None) with decode(q) == q @ A + shift, or None."""
        if not self.is_affine():
            return None
        layer = self.layers[0]
        A = (layer.matrix() / self.div.unsqueeze(1)).unsqueeze(0)
        return A, layer.shift

    def quantize(self, weight):
        """Latents as the decoder sees them: SGA mix or straight-through round."""
        if self.use_sga:
            return sga_quantize(weight, self.temperature, self.diff_sampling, getattr(self, "sga_uniforms", None))
        return StraightThrough.apply(weight)

    def forward(self, weight):
        out = self.layers(self.quantize(weight) / self.div)
        out = self.final_activation(out)
        if self.clamp_weights > 0.0:
            out = torch.clamp(out, min=-self.clamp_weights, max=self.clamp_weights)
        return out


class DecoderIdentity(nn.Module):
    """Pass-through with the decoder's interface. basic_latent_decoder.py:202-228."""

    def __init__(self):
        super().__init__()
        self.latent_dim = 1
        self.num_layers_dec = 0
        self.shift = False
        self.norm = "none"

    def reset_parameters(self, init_type, param=1.0):
        return

    def forward(self, x):
        return x

    def scale_norm(self):
        return 1

    def scale_grad_norm(self):
        return 1

    def size(self, use_torchac=False):
        return 0


class HierarchicalLatentDecoder(nn.Module):
    """One LatentDecoder per level (hierarchical_latent_decoder.py:3-15).

    `offsets` are the level boundaries in table rows. The reference builds them as
    cat(first_idx, lod_sizes[-1:]) (latent_grid.py:182), whose last entry is the last
    level's SIZE rather than the table length, so its last level is never decoded and keeps
    `torch.empty` garbage. That cannot be reproduced meaningfully; here the last
    boundary is the table length (documented deviation)."""

    def __init__(self, num_decoders, offsets, conf_decoder):
        super().__init__()
        self.num_decoders = num_decoders
        self.decoders = nn.ModuleList([LatentDecoder(**conf_decoder) for _ in range(num_decoders)])
        self.offsets = [int(o) for o in offsets]

    def forward(self, x):
        out = x.new_empty((x.size(0), self.decoders[0].channels))
        for l, dec in enumerate(self.decoders):
            a, b = self.offsets[l], self.offsets[l + 1]
            out[a:b] = dec(x[a:b])
        return out

    @property
    def temperature(self):
        return self.decoders[0].temperature

    @temperature.setter
    def temperature(self, value):
        for dec in self.decoders:
            dec.temperature = value

    @property
    def use_sga(self):
        return self.decoders[0].use_sga

    @use_sga.setter
    def use_sga(self, value):
        for dec in self.decoders:
            dec.use_sga = value

    @property
    def diff_sampling(self):
        return self.decoders[0].diff_sampling

    def size(self, use_torchac=False):
        return sum(p.numel() * torch.finfo(p.dtype).bits for p in self.parameters())

    def is_affine(self):
        return all(d.is_affine() for d in self.decoders)

    def affine_map(self):
        if not self.is_affine():
            return None
        maps = [d.affine_map() for d in self.decoders]
        A = torch.cat([m[0] for m in maps], dim=0)
        shift = torch.cat([m[1] for m in maps], dim=0) if maps[0][1] is not None else None
        return A, shift

    def quantize(self, weight):
        if self.use_sga:
            return sga_quantize(weight, self.temperature, self.diff_sampling)
        return StraightThrough.apply(weight)


class StraightThroughOneHot(torch.autograd.Function):
    """argmax one-hot over dim 0 forward, identity backward. multi_latent_decoder.py:15-25."""

    @staticmethod
    def forward(ctx, x):
        hot = torch.nn.functional.one_hot(torch.argmax(x, dim=0), num_classes=x.size(0))
        return hot.permute(1, 0).to(x)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output


class MultiLatentDecoderLayer(nn.Module):
    """Mixture of `num_decoders` linear maps selected per table row by alpha
    (multi_latent_decoder.py:27-77). Note the reference names the bias `use_shift`."""

    def __init__(self, in_features, out_features, ldecode_matrix, num_decoders=1, bias=False):
        super().__init__()
        self.in_features, self.out_features, self.ldecode_matrix = in_features, out_features, ldecode_matrix
        if "dft" in ldecode_matrix:
            self.dft = nn.Parameter(get_dft_matrix(in_features, out_features), requires_grad=False)
            self.scale = nn.Parameter(torch.empty((num_decoders, 1, out_features)))
        else:
            self.scale = nn.Parameter(torch.empty((num_decoders, in_features, out_features)))
        if bias:
            self.use_shift = nn.Parameter(torch.empty(num_decoders, 1, out_features))
        else:
            self.register_parameter("use_shift", None)
        if ldecode_matrix == "dft_fixed":
            self.scale.requires_grad_(False)

    def reset_parameters(self, param=1.0, init_type="normal"):
        if init_type == "normal":
            nn.init.normal_(self.scale, std=param)
        elif init_type == "uniform":
            nn.init.uniform_(self.scale, -param, param)
        elif init_type == "constant":
            nn.init.constant_(self.scale, val=param)
        if self.use_shift is not None:
            nn.init.zeros_(self.use_shift)

    def clamp(self, val=0.5):
        with torch.no_grad():
            self.scale.clamp_(-val, val)

    def forward(self, x, alpha):
        bias = self.use_shift if self.use_shift is not None else 0
        if "dft" in self.ldecode_matrix:
            w_out = torch.matmul(x, self.dft).unsqueeze(0) * self.scale + bias
        else:
            # the reference mixes with alpha once here and once more below (:68,:70)
            per_dec = torch.stack([torch.matmul(x, self.scale[i]) for i in range(self.scale.size(0))])
            w_out = torch.sum(per_dec * alpha.unsqueeze(-1), dim=0) + bias
        return torch.sum(w_out * alpha.unsqueeze(-1), dim=0)


class _MultiSequential(nn.Sequential):
    def forward(self, x, alpha):
        for module in self._modules.values():
            x = module(x, alpha) if isinstance(module, MultiLatentDecoderLayer) else module(x)
        return x


class MultiLatentDecoder(nn.Module):
    """multi_latent_decoder.py:84-210. Row-dependent (alpha) decode: not affine per level, so it
    always runs table-side in PyTorch ahead of the plain interpolation kernel."""

    def __init__(self, latent_dim, feature_dim, norm, ldecode_matrix, use_shift, num_entries, num_layers_dec=0,
                 hidden_dim_dec=0, activation="none", final_activation="none", clamp_weights=0.0, ldec_std=1.0,
                 num_decoders=1, alpha_std=1.0, use_sga=False, **kwargs):
        super().__init__()
        latent_dim = feature_dim if latent_dim == 0 else latent_dim
        self.ldecode_matrix, self.channels, self.latent_dim, self.norm = ldecode_matrix, feature_dim, latent_dim, norm
        self.div = nn.Parameter(torch.ones(latent_dim), requires_grad=False)
        self.num_layers_dec = num_layers_dec
        if num_layers_dec > 0:
            if hidden_dim_dec == 0:
                hidden_dim_dec = feature_dim
            if not isinstance(hidden_dim_dec, (tuple, list)):
                hidden_dim_dec = (hidden_dim_dec,) * num_layers_dec
            self.hidden_dim_dec = tuple(hidden_dim_dec)
        self.use_shift = use_shift
        self.act = _activation(activation)
        self.final_activation = _activation(final_activation)
        self.clamp_weights = clamp_weights
        self.num_decoders = num_decoders
        layers, width = [], latent_dim
        for l in range(num_layers_dec):
            hidden = self.hidden_dim_dec[l] or width
            layers += [MultiLatentDecoderLayer(width, hidden, ldecode_matrix, num_decoders, bias=use_shift), self.act]
            width = hidden
        layers.append(MultiLatentDecoderLayer(width, feature_dim, ldecode_matrix, num_decoders, bias=use_shift))
        self.alpha = nn.Parameter(torch.randn(num_decoders, num_entries) * alpha_std, requires_grad=True)
        self.temperature = 1.0
        self.layers = _MultiSequential(*layers)
        self.reset_parameters("normal", ldec_std)
        self.straight_through = True
        self.use_sga = use_sga
        self.diff_sampling = False

    def reset_parameters(self, init_type, param=0.5):
        for layer in self.layers.children():
            if isinstance(layer, MultiLatentDecoderLayer):
                layer.reset_parameters(param, init_type)

    def get_scale(self):
        assert self.num_layers_dec == 0, "Can only get scale for 0 hidden layers decoder!"
        return self.layers[0].scale

    def clamp(self, val=0.2):
        for layer in self.layers.children():
            if isinstance(layer, MultiLatentDecoderLayer):
                layer.clamp(val)

    def is_affine(self):
        return False

    def affine_map(self):
        return None

    def size(self, use_torchac=False):
        from . import bitstream
        fp_bits = sum(p.numel() * torch.finfo(p.dtype).bits for n, p in self.named_parameters() if "alpha" not in n)
        choice = torch.argmax(self.alpha, dim=0)
        _, counts = torch.unique(choice, return_counts=True)
        if not use_torchac:
            probs = counts / torch.sum(counts)
            info = torch.clamp(-1.0 * torch.log(probs + 1e-10) / math.log(2.0), 0, 1000)
            return torch.sum(info * counts).item() + fp_bits
        return bitstream.coded_bits(choice - choice.min()) + fp_bits

    def forward(self, weight):
        alpha = nn.functional.softmax(self.alpha / self.temperature, dim=0)
        if self.straight_through:
            alpha = StraightThroughOneHot.apply(alpha)
        q = sga_quantize(weight, self.temperature, self.diff_sampling) if self.use_sga else StraightThrough.apply(weight)
        out = self.final_activation(self.layers(q / self.div, alpha))
        if self.clamp_weights > 0.0:
            out = torch.clamp(out, min=-self.clamp_weights, max=self.clamp_weights)
        return out
