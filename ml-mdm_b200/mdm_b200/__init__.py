"""mdm_b200: Blackwell-native Matryoshka denoising path behind the reference's own interfaces.

    from mdm_b200.models import UNet, NestedUNet          # ml_mdm.models.unet / nested_unet
    from mdm_b200.diffusion import Diffusion, NestedDiffusion
    import mdm_b200.plugin; mdm_b200.plugin.register()    # overwrite ml_mdm.config registries
"""
__version__ = "0.1.0"
