"""Plug the B200 path into the reference's own registries (ml_mdm/config.py:9-63).

    import mdm_b200.plugin
    mdm_b200.plugin.register()        # after `import ml_mdm.models, ml_mdm.diffusion`

`get_model(name)` / `get_pipeline(name)` of the reference then return the classes of this package for
every architecture name (unet, nested_unet, nested2_unet, ...), so configs/models/*.yaml,
clis/train_parallel.py, clis/generate_batch.py and clis/generate_sample.py run unchanged: they build
`get_model(args.model)(3, 3, args.unet_config)` and `get_pipeline(args.model)(model, args.diffusion_config)`
(train_parallel.py:66-72).
"""


def register(config_module=None, parallel_names=False):
    """Overwrite MODEL_REGISTRY / PIPELINE_REGISTRY entries. With parallel_names=True the classes are
    registered additionally under '<name>_b200' instead of replacing the reference's."""
    if config_module is None:
        from ml_mdm import config as config_module  # the reference package must be importable
    from .diffusion import Diffusion, NestedDiffusion
    from .models import NestedUNet, UNet

    table = {"unet": (UNet, Diffusion), "nested_unet": (NestedUNet, NestedDiffusion)}
    for name, (model_cls, pipe_cls) in table.items():
        key = name + "_b200" if parallel_names else name
        config_module.MODEL_REGISTRY[key] = model_cls
        config_module.PIPELINE_REGISTRY[key] = pipe_cls
        if parallel_names:
            # arch -> {"model": registry key, "config": dataclass}: reuse the reference's config classes
            for arch, entry in list(config_module.MODEL_CONFIG_REGISTRY.items()):
                if entry["model"] == name and not arch.endswith("_b200"):
                    config_module.MODEL_CONFIG_REGISTRY[arch + "_b200"] = {"model": key, "config": entry["config"]}
            if name in config_module.PIPELINE_CONFIG_REGISTRY:
                config_module.PIPELINE_CONFIG_REGISTRY[key] = config_module.PIPELINE_CONFIG_REGISTRY[name]
    return config_module
