"""Configuration schema of the denoising path.

Field names, defaults and string-list parsing are those of the reference dataclasses so that its
YAML files (configs/models/cc12m_*.yaml) load unchanged:
  ResNetConfig / UNetConfig     ml_mdm/models/unet.py:44-156
  NestedUNetConfig (+2/3/4)     ml_mdm/models/nested_unet.py:21-75
  SamplerConfig                 ml_mdm/samplers.py:64-119
  DiffusionConfig / Nested...   ml_mdm/diffusion.py:29-50, 214-248
The model/pipeline classes of this package accept either these objects or the reference's own
config objects (they only read attributes).
"""
import dataclasses
import enum
import typing
from dataclasses import dataclass, field
from typing import Optional


def _ints(v, n=None):
    if v is None:
        return []
    if isinstance(v, str):
        v = [int(x) for x in v.split(",")] if v else []
    v = [int(x) for x in v]
    if n is not None and len(v) == 1:
        v = v * n
    return v


@dataclass
class ResNetConfig:
    num_channels: int = -1
    output_channels: int = -1
    num_groups_norm: int = 32
    dropout: float = 0.0
    use_attention_ffn: bool = False


@dataclass
class UNetConfig:
    num_resnets_per_resolution: typing.Any = "2"
    temporal_dim: Optional[int] = None
    attention_levels: typing.Any = "2,3"
    num_attention_layers: typing.Any = "1"
    num_temporal_attention_layers: typing.Any = None
    conditioning_feature_dim: int = -1
    conditioning_feature_proj_dim: int = -1
    num_lm_head_layers: int = 0
    masked_cross_attention: int = 1
    resolution_channels: typing.Any = "128,256,256,512,1024"
    skip_mid_blocks: bool = False
    skip_cond_emb: bool = False
    nesting: bool = False
    micro_conditioning: Optional[str] = None
    temporal_mode: bool = False
    temporal_spatial_ds: bool = False
    temporal_positional_encoding: bool = False
    resnet_config: ResNetConfig = field(default_factory=ResNetConfig)

    def __post_init__(self):
        self.resolution_channels = _ints(self.resolution_channels)
        n = len(self.resolution_channels)
        self.attention_levels = _ints(self.attention_levels)
        self.num_attention_layers = _ints(self.num_attention_layers, n)
        self.num_resnets_per_resolution = _ints(self.num_resnets_per_resolution, n)
        assert len(self.num_attention_layers) == n and len(self.num_resnets_per_resolution) == n
        if isinstance(self.resnet_config, dict):
            self.resnet_config = ResNetConfig(**self.resnet_config)


@dataclass
class NestedUNetConfig(UNetConfig):
    inner_config: typing.Any = field(default_factory=lambda: UNetConfig(nesting=True))
    skip_mid_blocks: bool = True
    skip_cond_emb: bool = True
    skip_inner_unet_input: bool = False
    skip_normalization: bool = False
    initialize_inner_with_pretrained: Optional[str] = None
    freeze_inner_unet: bool = False
    interp_conditioning: bool = False

    def __post_init__(self):
        super().__post_init__()
        if isinstance(self.inner_config, dict):
            self.inner_config = unet_config_from_dict(self.inner_config)
        if self.initialize_inner_with_pretrained == "None":
            self.initialize_inner_with_pretrained = None


# nested2/3/4 differ only in how deep inner_config nests (nested_unet.py:54-75)
Nested2UNetConfig = NestedUNetConfig
Nested3UNetConfig = NestedUNetConfig
Nested4UNetConfig = NestedUNetConfig


def unet_config_from_dict(d: dict):
    """dict (e.g. yaml['unet_config']) -> UNetConfig / NestedUNetConfig (by presence of inner_config)."""
    d = dict(d)
    cls = NestedUNetConfig if "inner_config" in d else UNetConfig
    names = {f.name for f in dataclasses.fields(cls)}
    kw = {k: (None if (isinstance(v, str) and v == "None") else v) for k, v in d.items() if k in names}
    return cls(**kw)


class _Type(enum.Enum):
    def __str__(self):
        return self.name.lower()

    __repr__ = __str__

    @classmethod
    def argparse(cls, s):
        try:
            return cls[str(s).upper()]
        except KeyError:
            return s


class ScheduleType(_Type):
    COSINE = 0
    DDPM = 2
    DEEPFLOYD = 3
    SIGMOID = 4


class PredictionType(_Type):
    DDPM = 3
    DDIM = 4
    V_PREDICTION = 5


class ThresholdType(_Type):
    NONE = 0
    CLIP = 1
    DYNAMIC = 2
    DYNAMIC_IF = 3


def _enum(cls, v):
    if v is None or isinstance(v, cls):
        return v
    name = getattr(v, "name", None)  # the reference's own enum members
    if name is not None:
        return cls[name]
    return cls.argparse(v)


@dataclass
class SamplerConfig:
    num_diffusion_steps: int = 32
    reproject_signal: bool = False
    schedule_type: typing.Any = ScheduleType.DDPM
    prediction_type: typing.Any = PredictionType.DDPM
    loss_target_type: typing.Any = None
    beta_start: float = 0.0001
    beta_end: float = 0.02
    threshold_function: typing.Any = ThresholdType.CLIP
    rescale_schedule: float = 1.0
    rescale_signal: Optional[float] = None
    schedule_shifted: bool = False
    schedule_shifted_power: float = 1

    def __post_init__(self):
        self.schedule_type = _enum(ScheduleType, self.schedule_type)
        self.prediction_type = _enum(PredictionType, self.prediction_type)
        self.loss_target_type = _enum(PredictionType, self.loss_target_type)
        self.threshold_function = _enum(ThresholdType, self.threshold_function)


@dataclass
class DiffusionConfig:
    sampler_config: typing.Any = field(default_factory=SamplerConfig)
    model_output_scale: float = 0
    use_vdm_loss_weights: bool = True

    def __post_init__(self):
        if isinstance(self.sampler_config, dict):
            names = {f.name for f in dataclasses.fields(SamplerConfig)}
            self.sampler_config = SamplerConfig(**{k: v for k, v in self.sampler_config.items() if k in names})


@dataclass
class NestedDiffusionConfig(DiffusionConfig):
    use_double_loss: bool = False
    multi_res_weights: Optional[str] = None
    no_use_residual: bool = False
    use_random_interp: bool = False
    mixed_ratio: Optional[str] = None
    random_downsample: bool = False
    average_downsample: bool = False
    mid_downsample: bool = False


def diffusion_config_from_dict(d: dict, nested: bool):
    cls = NestedDiffusionConfig if nested else DiffusionConfig
    names = {f.name for f in dataclasses.fields(cls)}
    kw = {k: (None if (isinstance(v, str) and v == "None") else v) for k, v in d.items() if k in names}
    return cls(**kw)


def load_yaml_configs(path: str, lm_dim: int = 2048):
    """Read a reference model YAML -> (unet_config, diffusion_config, is_nested). The language-model
    width is injected as train_parallel.py:65 does."""
    import yaml

    with open(path) as f:
        y = yaml.safe_load(f)
    ucfg = unet_config_from_dict(y["unet_config"])
    ucfg.conditioning_feature_dim = lm_dim
    c = ucfg
    while c is not None:  # pretrained-inner download needs the network; callers load checkpoints explicitly
        if hasattr(c, "initialize_inner_with_pretrained"):
            c.initialize_inner_with_pretrained = None
        c = getattr(c, "inner_config", None)
    nested = isinstance(ucfg, NestedUNetConfig)
    dcfg = diffusion_config_from_dict(y["diffusion_config"], nested)
    return ucfg, dcfg, nested
