"""Architecture registry + builders (drop-in for ``connectomics.models.architectures``,
``__init__.py:41-65`` there).  Importing this package registers ``mednext``, ``mednext_custom`` and ``monai_unet``."""

from .base import ConnectomicsModel
from .registry import (get_architecture_builder, get_architecture_info, is_architecture_available,
                       list_architectures, register_architecture, unregister_architecture)
from .mednext import (MedNeXt, MedNeXtBlock, MedNeXtMultiHeadWrapper, MedNeXtTaskHead, MedNeXtWrapper,
                      build_mednext, build_mednext_custom, create_mednext_v1)
from .monai_unet import MONAIModelWrapper, UNet as MonaiUNet, build_monai_unet
from .build import build_model


def print_available_architectures() -> None:
    for name, info in sorted(get_architecture_info().items()):
        print(f"{name:20s} {info['doc'].splitlines()[0]}")


__all__ = ["ConnectomicsModel", "register_architecture", "get_architecture_builder", "list_architectures",
           "is_architecture_available", "unregister_architecture", "get_architecture_info",
           "print_available_architectures", "build_model", "MedNeXt", "MedNeXtBlock", "MedNeXtWrapper",
           "MedNeXtTaskHead", "MedNeXtMultiHeadWrapper", "build_mednext", "build_mednext_custom",
           "create_mednext_v1", "MONAIModelWrapper", "MonaiUNet", "build_monai_unet"]
