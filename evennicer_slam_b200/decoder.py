"""Drop-in decoders: same class names, constructor arguments, attribute names and ``state_dict``
keys as ``src/conv_onet/models/decoder.py`` of the reference, so ``load_pretrain``
(EvenNICER_SLAM.py:184-215), ``Logger.log``, ``copy.deepcopy``, ``.to()`` and
``.share_memory()`` keep working.  The arithmetic is NOT here: ``NICE.forward`` calls the fused
CUDA decode (``ens_eval_points``); the differentiable path is ``Renderer.render_batch_ray``.

  reference class                     here
  GaussianFourierFeatureTransform     decoder.py:7-30    parameter ``_B`` (3,93) ~ N(0, 25^2)
  DenseLayer                          decoder.py:70-79   xavier_uniform(gain(activation)), zero bias
  MLP                                 decoder.py:91-203  fc_c.{i}, embedder._B, pts_linears.{i}, output_linear
  MLP_no_xyz                          decoder.py:206-274 pts_linears.{i}, output_linear
  NICE                                decoder.py:277-342 {coarse,middle,fine,color}_decoder, forward(p, c_grid, stage)
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .scene import SceneCache


class GaussianFourierFeatureTransform(nn.Module):
    def __init__(self, num_input_channels, mapping_size=93, scale=25, learnable=True):
        super().__init__()
        B = torch.randn((num_input_channels, mapping_size)) * scale
        if learnable:
            self._B = nn.Parameter(B)
        else:
            self._B = B


class DenseLayer(nn.Linear):
    def __init__(self, in_dim: int, out_dim: int, activation: str = "relu", *args, **kwargs) -> None:
        self.activation = activation
        super().__init__(in_dim, out_dim, *args, **kwargs)

    def reset_parameters(self) -> None:
        nn.init.xavier_uniform_(self.weight, gain=nn.init.calculate_gain(self.activation))
        if self.bias is not None:
            nn.init.zeros_(self.bias)


def _unsupported(what):
    raise NotImplementedError(
        f"{what} is outside the fused NICE hot path (hidden 32, c_dim 32, fourier embedding, skips=[2], "
        "5 blocks); the B200 kernels implement exactly the shipped NICE configuration")


class MLP(nn.Module):
    """Parameter container for one Fourier-feature decoder (decoder.py:91-166)."""

    def __init__(self, name='', dim=3, c_dim=128, hidden_size=256, n_blocks=5, leaky=False, sample_mode='bilinear',
                 color=False, skips=[2], grid_len=0.16, pos_embedding_method='fourier', concat_feature=False):
        super().__init__()
        if pos_embedding_method != 'fourier' or hidden_size != 32 or n_blocks != 5 or list(skips) != [2] \
                or leaky or sample_mode != 'bilinear' or dim != 3 or c_dim not in (32, 64):
            _unsupported(f"MLP(name={name!r}, c_dim={c_dim}, hidden_size={hidden_size}, n_blocks={n_blocks}, "
                         f"skips={skips}, pos_embedding_method={pos_embedding_method!r})")
        self.name = name
        self.color = color
        self.no_grad_feature = False
        self.c_dim = c_dim
        self.grid_len = grid_len
        self.concat_feature = concat_feature
        self.n_blocks = n_blocks
        self.skips = skips
        self.sample_mode = sample_mode
        embedding_size = 93
        self.fc_c = nn.ModuleList([nn.Linear(c_dim, hidden_size) for _ in range(n_blocks)])
        self.embedder = GaussianFourierFeatureTransform(dim, mapping_size=embedding_size, scale=25)
        self.pts_linears = nn.ModuleList(
            [DenseLayer(embedding_size, hidden_size, activation="relu")] +
            [DenseLayer(hidden_size, hidden_size, activation="relu") if i not in self.skips
             else DenseLayer(hidden_size + embedding_size, hidden_size, activation="relu")
             for i in range(n_blocks - 1)])
        self.output_linear = DenseLayer(hidden_size, 4 if color else 1, activation="linear")

    def forward(self, p, c_grid=None):
        raise NotImplementedError(
            "single-decoder calls are not part of the fused path; call NICE.forward(p, c_grid, stage) "
            "or Renderer.render_batch_ray")


class MLP_no_xyz(nn.Module):
    """Parameter container for the coarse decoder (decoder.py:206-252)."""

    def __init__(self, name='', dim=3, c_dim=128, hidden_size=256, n_blocks=5, leaky=False, sample_mode='bilinear',
                 color=False, skips=[2], grid_len=0.16):
        super().__init__()
        if hidden_size != 32 or n_blocks != 5 or list(skips) != [2] or leaky or c_dim != 32 or color:
            _unsupported(f"MLP_no_xyz(c_dim={c_dim}, hidden_size={hidden_size}, n_blocks={n_blocks})")
        self.name = name
        self.no_grad_feature = False
        self.color = color
        self.grid_len = grid_len
        self.c_dim = c_dim
        self.n_blocks = n_blocks
        self.skips = skips
        self.sample_mode = sample_mode
        self.pts_linears = nn.ModuleList(
            [DenseLayer(hidden_size, hidden_size, activation="relu")] +
            [DenseLayer(hidden_size, hidden_size, activation="relu") if i not in self.skips
             else DenseLayer(hidden_size + c_dim, hidden_size, activation="relu") for i in range(n_blocks - 1)])
        self.output_linear = DenseLayer(hidden_size, 1, activation="linear")

    def forward(self, p, c_grid, **kwargs):
        raise NotImplementedError(
            "single-decoder calls are not part of the fused path; call NICE.forward(p, c_grid, stage='coarse')")


class NICE(nn.Module):
    """Neural Implicit Scalable Encoding, fused (decoder.py:277-342)."""

    def __init__(self, dim=3, c_dim=32, coarse_grid_len=2.0, middle_grid_len=0.16, fine_grid_len=0.16,
                 color_grid_len=0.16, hidden_size=32, coarse=False, pos_embedding_method='fourier'):
        super().__init__()
        if coarse:
            self.coarse_decoder = MLP_no_xyz(name='coarse', dim=dim, c_dim=c_dim, color=False,
                                             hidden_size=hidden_size, grid_len=coarse_grid_len)
        self.middle_decoder = MLP(name='middle', dim=dim, c_dim=c_dim, color=False, skips=[2], n_blocks=5,
                                  hidden_size=hidden_size, grid_len=middle_grid_len,
                                  pos_embedding_method=pos_embedding_method)
        self.fine_decoder = MLP(name='fine', dim=dim, c_dim=c_dim * 2, color=False, skips=[2], n_blocks=5,
                                hidden_size=hidden_size, grid_len=fine_grid_len, concat_feature=True,
                                pos_embedding_method=pos_embedding_method)
        self.color_decoder = MLP(name='color', dim=dim, c_dim=c_dim, color=True, skips=[2], n_blocks=5,
                                 hidden_size=hidden_size, grid_len=color_grid_len,
                                 pos_embedding_method=pos_embedding_method)
        self._ens_cache = None

    def __deepcopy__(self, memo):
        # the native-layout cache is per-object state, not model state (Tracker deep-copies the decoders)
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k == "_ens_cache" else copy.deepcopy(v, memo)
        return new

    def forward(self, p, c_grid, stage='middle', **kwargs):
        """Output occupancy/color in different stages: (N,4) float32 [r,g,b,occ]; no bound mask."""
        from .functional import RenderSetup, eval_points
        # Forward-only: Renderer.eval_points is the only differentiable caller in the reference
        # (Renderer.py:51) and the fused Renderer does not route through here; Mesher.eval_points
        # (Mesher.py:308) runs under no_grad.
        if self._ens_cache is None:
            self._ens_cache = SceneCache()
        bound = self.bound
        cbound = self.coarse_decoder.bound if hasattr(self, "coarse_decoder") else bound
        setup = RenderSetup(stage, 0, 0, None, None, bound, cbound, self._ens_cache)
        with torch.no_grad():
            return eval_points(setup, c_grid, self, p, apply_mask=False)


decoder_dict = {'nice': NICE}
