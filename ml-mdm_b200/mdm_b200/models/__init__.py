from .unet import UNet  # noqa: F401
from .nested_unet import NestedUNet  # noqa: F401
