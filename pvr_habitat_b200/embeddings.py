"""Drop-in for the reference's src/embeddings.py (EmbeddingNet / _get_embedding / UberModel / EmbeddingWrapper).

Same constructor arguments, attributes (`embedding_name, in_channels, embedding, transforms, in_shape, out_size,
device, training`), state_dict keys (`embedding.*`, empty for uber models — reference quirk D8, src/embeddings.py:45-53)
and return types (`numpy (N, O)` float32 in eval mode, squeezed when N == 1, src/embeddings.py:398-402). The arithmetic
runs in libpvr_b200: the fused uint8 preprocessing kernel and the tcgen05 encoder program. There is no CPU path.
"""
import ctypes
import os

import numpy as np
import torch
from torch import nn

from . import _lib
from . import program as prg
from .vision_models import clip_rn
from .vision_models import clip_vit
from .vision_models import mae as mae_vit
from .vision_models import maskrcnn
from .vision_models.moco import moco_conv3_compressed, moco_conv4_compressed, moco_conv5, random_init_allowed
from .vision_models.resnet import resnet_conv3_compressed, resnet_conv4_compressed, resnet_conv5
from .vision_models.resnet_params import ResNet50Params, ResNetBasicParams

IMAGENET_MEAN, IMAGENET_STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
CLIP_MEAN, CLIP_STD = [0.48145466, 0.4578275, 0.40821073], [0.26862954, 0.26130258, 0.27577711]


def resize_geometry(h, w, size=256, crop=224):
    """torchvision Resize(int) + CenterCrop geometry (transforms/functional.py:368-384, 592-594)."""
    if h <= w:
        rh, rw = size, int(size * w / h)
    else:
        rh, rw = int(size * h / w), size
    top = int(round((rh - crop) / 2.0))
    left = int(round((rw - crop) / 2.0))
    return rh, rw, top, left


class Transforms(nn.Module):
    """Resize(256) -> CenterCrop(224) -> ConvertImageDtype(float) -> Normalize (src/embeddings.py:80-85) as ONE kernel.

    Calling it on an NCHW uint8 tensor (the reference's calling convention, src/embeddings.py:393) returns the
    NCHW float32 tensor the reference's nn.Sequential of torchvision transforms returns, bit for bit.
    """

    def __init__(self, mean=IMAGENET_MEAN, std=IMAGENET_STD, size=256, crop=224, interpolation='bilinear'):
        super().__init__()
        self.mean, self.std, self.size, self.crop = list(mean), list(std), size, crop
        assert interpolation in ('bilinear', 'bicubic', 'bicubic_aa', 'bilinear_float_rows02')
        # 'bicubic': T.Resize(256, interpolation=3) of the MAE encoders; 'bicubic_aa': CLIP's antialiased bicubic Resize
        # (pvr_preprocess_u8_aa, csrc/preprocess_aa.cu); 'bilinear_float_rows02': the maskrcnn_l3 transforms
        # (src/embeddings.py:283-294) — `_rgb_to_bgr` (which permutes ROWS 0 and 2, not channels), `.float()`, Resize of
        # the float image (no uint8 rounding), CenterCrop, Normalize on the 0..255 scale
        self.interpolation = interpolation
        self.identity_resize_only = False  # debugging aid: reject frames that would need an actual resize

    def format_bytes_per_frame(self, fmt, h=224, w=224):
        """HBM bytes one frame costs in output format `fmt` (uint8 frame read + formatted frame written)."""
        c = self.crop
        out = {_lib.PVR_FMT_NCHW_F32: 3 * c * c * 4, _lib.PVR_FMT_NHWC4_BF16: c * c * 4 * 2,
               _lib.PVR_FMT_STEM_BF16: c * (c // 2) * 32 * 2, _lib.PVR_FMT_NHWC4_F32: c * c * 4 * 4,
               _lib.PVR_FMT_STEM_PAD_BF16: c * (c + 8) * 4 * 2}[fmt & 0xff]
        return h * w * 3 + out

    def run(self, obs_nhwc_u8, n_frames, out_ptr, fmt, sample_major):
        """obs: CUDA uint8 (N, H, W, 3*n_frames) contiguous; writes n_frames*N images at `out_ptr`."""
        n, h, w, ch = obs_nhwc_u8.shape
        assert ch == 3 * n_frames and obs_nhwc_u8.dtype == torch.uint8 and obs_nhwc_u8.is_contiguous()
        if self.identity_resize_only and (h != self.size or w != self.size):
            raise NotImplementedError(f"CLIP preprocessing of {h}x{w} frames needs the antialiased bicubic resize; "
                                      f"only {self.size}x{self.size} frames (identity resize) are supported")
        rh, rw, top, left = resize_geometry(h, w, self.size, self.crop)
        mean = (ctypes.c_float * 3)(*self.mean)
        std = (ctypes.c_float * 3)(*self.std)
        fn, name = _lib.lib().pvr_preprocess_u8, "pvr_preprocess_u8"
        if self.interpolation == 'bicubic':
            fmt |= _lib.PVR_RESIZE_BICUBIC
        elif self.interpolation == 'bilinear_float_rows02':
            fmt |= _lib.PVR_RESIZE_FLOAT | _lib.PVR_SWAP_ROWS_0_2
        elif self.interpolation == 'bicubic_aa' and (rh, rw) != (h, w):
            # (torchvision leaves an image whose short side already has the requested size untouched,
            # tv:transforms/functional.py:468-471: that case is the scale-1 path of the bilinear kernel = a copy)
            fn, name = _lib.lib().pvr_preprocess_u8_aa, "pvr_preprocess_u8_aa"
        with torch.cuda.device(obs_nhwc_u8.device):
            _lib.check(fn(obs_nhwc_u8.data_ptr(), n, h, w, n_frames, rh, rw, top, left, self.crop, mean, std, out_ptr,
                          fmt, int(sample_major), _lib.current_stream_ptr()), name)

    def forward(self, x):
        if not x.is_cuda:
            raise _lib.PvrError("pvr_habitat_b200 transforms run on CUDA only (no CPU fallback)")
        if x.dtype != torch.uint8:
            raise _lib.PvrError("transforms expect the uint8 frames the reference feeds them")
        nhwc = x.permute(0, 2, 3, 1).contiguous()
        out = torch.empty(x.shape[0], 3, self.crop, self.crop, dtype=torch.float32, device=x.device)
        self.run(nhwc, 1, out.data_ptr(), _lib.PVR_FMT_NCHW_F32, False)
        return out


def init(module, weight_init, bias_init, gain=1):
    weight_init(module.weight.data, gain=gain)
    bias_init(module.bias.data)
    return module


class SmallConvParams(nn.Sequential):
    """Parameter container of the 'random' PVR; the arithmetic is program.add_small_conv (no torch forward)."""
    variant = 'small_conv'

    def __init__(self, in_channels=3):
        init_ = lambda m: init(m, nn.init.orthogonal_, lambda x: nn.init.constant_(x, 0),  # noqa: E731
                               nn.init.calculate_gain('relu'))
        layers = []
        for i in range(5):
            layers += [init_(nn.Conv2d(in_channels if i == 0 else 32, 32, kernel_size=(3, 3), stride=2, padding=1)),
                       nn.ELU()]
        super().__init__(*layers)
        self.out_size = 32 * 7 * 7  # 224 -> 112 -> 56 -> 28 -> 14 -> 7

    def forward(self, x):
        raise _lib.PvrError("SmallConvParams holds parameters only; use EmbeddingNet (CUDA program), no torch fallback")


class UberModel(nn.Module):
    """Concatenation of several encoders (src/embeddings.py:44-57). `models` is a plain list, as in the reference,
    so an uber EmbeddingNet has an empty state_dict."""

    def __init__(self, models):
        super(UberModel, self).__init__()
        self.models = models
        assert all(models[0].training == m.training for m in models)
        self.training = models[0].training
        self.out_size = sum(m.out_size for m in models)

    def to(self, device):
        self.models = [m.to(device=device) for m in self.models]
        return self


_MOCO_CONV5 = {
    'demy': 'demy.pth', 'moco_aug': 'moco_aug.pth.tar', 'moco_aug_habitat': 'moco_aug_habitat_64.pth',
    'moco_aug_mujoco': 'moco_aug_mujoco.pth', 'moco_aug_uber': 'moco_aug_uber.pth',
    'moco_aug_places': 'moco_aug_places.pth.tar', 'moco_croponly': 'moco_croponly.pth',
    'moco_croponly_places': 'moco_croponly_places.pth', 'moco_croponly_habitat': 'moco_croponly_habitat_64.pth',
    'moco_croponly_mujoco': 'moco_croponly_mujoco.pth', 'moco_croponly_uber': 'moco_croponly_uber.pth',
    'moco_coloronly': 'moco_coloronly.pth',
}
_MOCO_L4 = {n: n + '.pth' for n in ('moco_aug_l4', 'moco_aug_places_l4', 'moco_croponly_l4',
                                    'moco_croponly_places_l4')}
_MOCO_L3 = {n: n + '.pth' for n in ('moco_aug_l3', 'moco_aug_places_l3', 'moco_croponly_l3',
                                    'moco_croponly_places_l3')}
_RESNET = {
    'resnet50_places': (resnet_conv5, 'resnet50_places.pth.tar'),
    'resnet50_l4': (resnet_conv4_compressed, 'resnet50_l4.pth.tar'),
    'resnet50_l3': (resnet_conv3_compressed, 'resnet50_l3.tar'),
    'resnet50_places_l4': (resnet_conv4_compressed, 'resnet50_places_l4.tar'),
    'resnet50_places_l3': (resnet_conv3_compressed, 'resnet50_places_l3.tar'),
}
_UBER_PARTS = {'3': '_l3', '4': '_l4', '5': ''}


# torchvision's ImageNet checkpoints (the files `pretrained=True` downloads under torchvision 0.10, the reference's pin)
_TORCHVISION_FILES = {'resnet18': 'resnet18-f37072fd.pth', 'resnet34': 'resnet34-b627a593.pth',
                      'resnet50': 'resnet50-0676ba61.pth'}


def _load_torchvision(model, name, pretrained):
    """`pretrained=True` in the reference downloads torchvision's ImageNet weights; there is no network here, so the
    file is looked up locally: $PVR_TORCHVISION_WEIGHTS/<file>, then torch.hub's checkpoint cache (where torchvision
    itself would have put it). A missing file raises — silently returning a random network would write a valid-looking
    dataset of noise embeddings — unless `allow_random_init()` is active (tests / benchmarks) or pretrained=False."""
    if not pretrained:
        return model
    fname = _TORCHVISION_FILES[name]
    dirs = [os.environ.get("PVR_TORCHVISION_WEIGHTS"), os.path.join(torch.hub.get_dir(), "checkpoints")]
    for d in filter(None, dirs):
        path = os.path.join(d, fname)
        if os.path.isfile(path):
            sd = {k: v for k, v in torch.load(path, map_location='cpu').items() if not k.startswith('fc.')}
            model.load_state_dict(sd, strict=True)
            return model
    if random_init_allowed():
        return model
    raise FileNotFoundError(
        f"EmbeddingNet('{name}', pretrained=True): torchvision's {fname} was not found in $PVR_TORCHVISION_WEIGHTS or "
        f"{dirs[1]} and cannot be downloaded offline; pass pretrained=False for a random-init network")


def _get_embedding(embedding_name='random', in_channels=3, pretrained=True, train=False):
    """Same names and return convention as src/embeddings.py:60-332: `(model, transforms)`."""
    transforms = Transforms(IMAGENET_MEAN, IMAGENET_STD)
    assert in_channels == 3, 'Current models accept 3-channel inputs only.'

    if embedding_name == 'random':
        # FIXED 5-LAYER CONV (src/embeddings.py:90-106): built with the same layer order / init calls as the
        # reference, so the same torch seed gives the same weights and the same state_dict keys ('0.weight', ...).
        model = SmallConvParams(in_channels)
    elif embedding_name in ('resnet18', 'resnet34'):
        # torchvision.models.resnet18 / resnet34(pretrained=...), fc -> Identity (src/embeddings.py:112-117)
        model = _load_torchvision(ResNetBasicParams(embedding_name), embedding_name, pretrained)
    elif embedding_name == 'resnet50':
        # torchvision.models.resnet50(pretrained=...), fc -> Identity (src/embeddings.py:118-120)
        model = _load_torchvision(ResNet50Params('conv5'), embedding_name, pretrained)
    elif embedding_name in _MOCO_CONV5:
        model = moco_conv5(checkpoint_path=_MOCO_CONV5[embedding_name])
    elif embedding_name in _MOCO_L4:
        model = moco_conv4_compressed(checkpoint_path=_MOCO_L4[embedding_name])
    elif embedding_name in _MOCO_L3:
        model = moco_conv3_compressed(checkpoint_path=_MOCO_L3[embedding_name])
    elif embedding_name in _RESNET:
        fn, path = _RESNET[embedding_name]
        model = fn(checkpoint_path=path)
    elif '_uber_' in embedding_name and embedding_name.startswith('moco_'):
        # e.g. moco_aug_places_uber_345 -> [moco_aug_places_l3, moco_aug_places_l4, moco_aug_places]
        base, taps = embedding_name.split('_uber_')
        if base not in ('moco_aug', 'moco_aug_places', 'moco_croponly', 'moco_croponly_places') or \
                taps not in ('345', '35', '34', '45'):
            raise NotImplementedError("Requested model not available.")
        model = UberModel([_get_embedding(base + _UBER_PARTS[t])[0] for t in taps])
    elif embedding_name in ('mae_base', 'mae_large', 'mae_huge'):
        # src/embeddings.py:137-148; MAE frames are resized with bicubic interpolation (src/embeddings.py:81)
        model = mae_vit.load(embedding_name)
        transforms = Transforms(IMAGENET_MEAN, IMAGENET_STD, interpolation='bicubic')
    elif 'clip' in embedding_name:
        # src/embeddings.py:298-314: clip.load("ViT-B/32") + CLIP's own normalisation. Resize(224, bicubic,
        # antialiased) + CenterCrop(224): the identity for 224x224 frames; any other size (Habitat renders 64x64,
        # habitat_config/nav_task.yaml:10-12) goes through pvr_preprocess_u8_aa (csrc/preprocess_aa.cu).
        if embedding_name == 'clip_vit':
            model, _ = clip_vit.load("ViT-B/32", device='cpu')
        elif embedding_name == 'clip_vit_b16':  # BASELINE configs[2] geometry (CLIP block structure, patch 16)
            model, _ = clip_vit.load("ViT-B/16", device='cpu')
        elif embedding_name == 'clip_rn50':  # src/embeddings.py:305-306
            model, _ = clip_rn.load("RN50", device='cpu')
        else:
            raise NotImplementedError("Requested model not available.")
        transforms = Transforms(CLIP_MEAN, CLIP_STD, size=model.visual.input_resolution,
                                crop=model.visual.input_resolution, interpolation='bicubic_aa')
    elif embedding_name == 'maskrcnn_l3':
        # src/embeddings.py:283-295: detectron2 R50-C4 backbone through res4 + the 1024 -> 11 compression block; the
        # frames stay on the 0..255 scale (Normalize with detectron2's pixel means, std 1) and are resized as floats
        model = maskrcnn.mask_rcnn_model(checkpoint_path='maskrcnn_l3.pth')
        transforms = Transforms(maskrcnn.PIXEL_MEAN, [1.0, 1.0, 1.0], interpolation='bilinear_float_rows02')
    elif embedding_name == 'true_state':
        return nn.Sequential(nn.Identity()), nn.Sequential(nn.Identity())
    else:
        raise NotImplementedError("Requested model not available.")

    if train:
        raise NotImplementedError("pvr_habitat_b200 runs the frozen-encoder path only (train=False).")
    model.eval()
    for p in model.parameters():
        p.requires_grad = False
    return model, transforms


def _program_sd(m, sd):
    """Containers whose keys are not torchvision's (maskrcnn.MaskRCNNBackboneParams: detectron2 names) translate
    their state_dict into the naming program.add_resnet50 reads."""
    return m.program_state_dict(sd) if hasattr(m, 'program_state_dict') else sd


def build_encoder(model, device, hw=224, precision='bf16'):
    """Compile `model` (ResNet50Params or UberModel of them) into one pvr_encoder program on `device`.

    precision: 'bf16' (tensor-core kernels) or 'fp32' (the north star's parity mode: float32 CUDA-core kernels,
    csrc/conv_f32.cu)."""
    prog = prg.Program()
    if precision == 'fp32':
        in_slot = prog.new_slot(hw * hw * 4 * 2)  # slot 0: NHWC4 float32 frames (two bf16 elements per float)
        parts = model.models if isinstance(model, UberModel) else [model]
        off = 0
        for m in parts:
            sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
            if isinstance(m, SmallConvParams):
                off += prg.add_small_conv_f32(prog, sd, in_slot, off, hw)
            elif isinstance(m, ResNetBasicParams):
                off += prg.add_resnet_basic_f32(prog, sd, m.LAYERS[m.name], in_slot, off, hw)
            else:
                off += prg.add_resnet50_f32(prog, _program_sd(m, sd), m.variant, in_slot, off, hw,
                                            stride_in_1x1=getattr(m, 'stride_in_1x1', False))
        prog.emb_width = off
        enc = prog.finish(device)
        enc.input_format = _lib.PVR_FMT_NHWC4_F32
        return enc
    if precision != 'bf16':
        raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
    if isinstance(model, SmallConvParams):
        in_slot = prog.new_slot(hw * hw * 4)  # slot 0: NHWC4 bf16 frames
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        prog.emb_width = prg.add_small_conv(prog, sd, in_slot, 0, hw)
        enc = prog.finish(device)
        enc.input_format = _lib.PVR_FMT_NHWC4_BF16
        return enc
    # slot 0: the stem's input from the preprocessing kernel. Default: padded NHWC4 rows (PVR_FMT_STEM_PAD_BF16,
    # (hw + 8) * 4 elements per row) that the stem's tensor map expands into its 8-column windows; PVR_STEM_EXPANDED=1
    # (A/B switch) or a frame size the patch-resident stem does not take: the materialised W-expanded layout.
    compact = hw % 32 == 0 and os.environ.get("PVR_STEM_EXPANDED") != "1"
    in_slot = prog.new_slot(hw * (hw + 8) * 4 if compact else hw * (hw // 2) * 32)
    parts = model.models if isinstance(model, UberModel) else [model]
    off = 0
    for m in parts:
        sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
        if isinstance(m, ResNetBasicParams):
            off += prg.add_resnet_basic(prog, sd, m.LAYERS[m.name], in_slot, off, hw, compact_stem=compact)
        else:
            off += prg.add_resnet50(prog, _program_sd(m, sd), m.variant, in_slot, off, hw, compact_stem=compact,
                                    stride_in_1x1=getattr(m, 'stride_in_1x1', False))
    prog.emb_width = off
    enc = prog.finish(device)
    enc.input_format = _lib.PVR_FMT_STEM_PAD_BF16 if compact else _lib.PVR_FMT_STEM_BF16
    return enc


class EmbeddingNet(nn.Module):
    """
    Input shape must be (N, H, W, 3), where N is the number of frames.
    The output shape will be (N, O), where O is the embedding size.  (src/embeddings.py:339-402)
    """

    def __init__(self, embedding_name, in_channels=3, pretrained=True, train=False, disable_cuda=False):
        super(EmbeddingNet, self).__init__()
        self.embedding_name = embedding_name
        if self.embedding_name == 'true_state':
            return
        self.in_channels = in_channels
        self.embedding, self.transforms = _get_embedding(embedding_name, in_channels, pretrained, train)
        self.in_shape = torch.Size([in_channels, self.transforms.crop, self.transforms.crop])
        self.out_size = int(self.embedding.out_size)
        if torch.cuda.is_available() and not disable_cuda:
            self.device = torch.device('cuda', torch.cuda.current_device())
        else:
            self.device = torch.device('cpu')
        self.embedding = self.embedding.to(device=self.device)
        self.training = self.embedding.training
        self._encoder = None
        self._emb = None
        self.max_images_per_pass = 1024
        self.precision = 'bf16'

    def set_precision(self, precision):
        """'bf16' (default, tensor cores) or 'fp32' (parity mode: float32 end to end on the CUDA cores, every encoder:
        ResNets, small conv, CLIP / MAE ViTs). The constructor signature stays the reference's, hence a setter."""
        if precision not in ('bf16', 'fp32'):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
        if precision != self.precision:
            self.precision = precision
            self._encoder = None
        return self

    # ---- weights changed -> recompile the program lazily
    # keys of openai/CLIP's text tower: part of the reference's `embedding_model_state_dict` for clip_* encoders (its
    # `embedding` is the whole CLIP model), not of the image path
    _CLIP_TEXT_KEYS = ("embedding.transformer.", "embedding.token_embedding.", "embedding.positional_embedding",
                       "embedding.ln_final.", "embedding.text_projection", "embedding.logit_scale",
                       "embedding.input_resolution", "embedding.context_length", "embedding.vocab_size")

    def load_state_dict(self, state_dict, *args, **kwargs):
        self._encoder = None
        if 'clip' in self.embedding_name:
            # A checkpoint written by the reference holds the full CLIP model; only `embedding.visual.*` exists here.
            # The text tower is dropped, anything else unknown still raises (strict). The other direction is one-way:
            # a checkpoint written here has no text tower, so the reference's strict load rejects it.
            state_dict = {k: v for k, v in state_dict.items() if not k.startswith(self._CLIP_TEXT_KEYS)}
        return super().load_state_dict(state_dict, *args, **kwargs)

    def invalidate(self):
        """Call after mutating `self.embedding` parameters in place."""
        self._encoder = None

    def _require_cuda(self):
        if self.device.type != 'cuda':
            raise _lib.PvrError("EmbeddingNet: CUDA device required — pvr_habitat_b200 has no CPU fallback "
                                "(disable_cuda=True is only meaningful for the reference implementation).")

    def encoder(self):
        self._require_cuda()
        if self._encoder is None:
            if isinstance(self.embedding, (clip_vit.CLIPImageModel, clip_rn.CLIPResNetModel, mae_vit.MAEParams)):
                self.embedding.invalidate()
                self._encoder = self.embedding.runner(self.device, self.precision)
            else:
                self._encoder = build_encoder(self.embedding, self.device, self.transforms.crop, self.precision)
        return self._encoder

    def preprocess(self, observation):
        """(N, H, W, 3) uint8 -> (N, 3, 224, 224) float32 CUDA tensor, bit-identical to the reference transforms."""
        self._require_cuda()
        obs = observation.to(device=self.device).contiguous()
        out = torch.empty(obs.shape[0], 3, self.transforms.crop, self.transforms.crop, dtype=torch.float32,
                          device=self.device)
        self.transforms.run(obs, 1, out.data_ptr(), _lib.PVR_FMT_NCHW_F32, False)
        return out

    def embed(self, observation, n_frames=1, out=None):
        """Fused path: (N, H, W, 3*n_frames) uint8 -> CUDA float32 (N, n_frames*O).

        Covers the host-side frame split / regroup of main_bc_1.py:128-136 and save_embedded_obs.py:149-155:
        frame f of sample i lands in out[i, f*O:(f+1)*O].
        """
        self._require_cuda()
        n = observation.shape[0]
        if out is None:
            out = torch.empty(n, n_frames * self.out_size, dtype=torch.float32, device=self.device)
        enc = self.encoder()
        # bound the activation workspace (ResNet-50: 6.8 MB per image): long observation arrays are embedded in
        # passes of at most `max_images_per_pass` images, each pass writing its rows of `out`
        step = max(1, self.max_images_per_pass // n_frames)
        for lo in range(0, n, step):
            obs = observation[lo:lo + step].to(device=self.device, non_blocking=True)
            if not obs.is_contiguous():
                obs = obs.contiguous()
            m = obs.shape[0]
            enc.bind(m * n_frames)
            self.transforms.run(obs, n_frames, enc.slot0, enc.input_format, True)
            enc.forward(out[lo:lo + m], self.out_size)
        return out

    def forward(self, observation):
        if self.embedding_name == 'true_state':
            return observation.squeeze().cpu().numpy()
        # observation.shape -> (N, H, W, 3)
        out = self.embed(observation, 1)
        return out.view(-1, self.out_size).squeeze().cpu().numpy()


try:  # gym is the simulator side's dependency; without it the adaptor keeps working on any object with the same duck type
    import gym
    from gym.spaces.box import Box
    _ObservationWrapper = gym.ObservationWrapper
except ImportError:
    class Box(object):
        """Minimal stand-in for gym.spaces.Box (shape / bounds only)."""

        def __init__(self, low=None, high=None, shape=None, dtype=None):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    class _ObservationWrapper(object):
        """Minimal stand-in for gym.ObservationWrapper: reset / step pass observations through `observation`."""

        def __init__(self, env):
            self.env = env
            self.observation_space = getattr(env, "observation_space", None)
            self.action_space = getattr(env, "action_space", None)

        def reset(self, **kwargs):
            return self.observation(self.env.reset(**kwargs))

        def step(self, action):
            observation, reward, done, info = self.env.step(action)
            return self.observation(observation), reward, done, info

        def __getattr__(self, name):
            return getattr(self.env, name)


class EmbeddingWrapper(_ObservationWrapper):
    """src/embeddings.py:409-444: (H, W, 3n) uint8 frames of one environment step -> flat (n*O,) float32 embedding
    (frame f in columns [f*O, (f+1)*O), `src/embeddings.py:441-444`). Batches of <= 8 images replay from a CUDA graph
    (pvr_encoder_forward), which is what makes the per-step latency of a rollout."""

    def __init__(self, env, embedding):
        _ObservationWrapper.__init__(self, env)
        in_channels = env.observation_space.shape[2]
        assert in_channels % 3 == 0, "Only RGB images are supported."
        self.in_channels = 3
        self.n_frames = in_channels // 3
        self.embedding = embedding
        self.observation_space = Box(low=-np.inf, high=np.inf,
                                     shape=(self.embedding.out_size * self.n_frames,))

    def observation(self, observation):
        obs = torch.from_numpy(np.ascontiguousarray(observation))[None]
        return self.embedding.embed(obs, self.n_frames).flatten().cpu().numpy()
