"""deepbedmap_b200 -- B200-native ESRGAN hot path of weiji14/deepbedmap (generator,
discriminator, training step, tiled continent predictor) behind the reference's call surface."""
from ._lib import DeepBedMapError, LIB_PATH  # noqa: F401

__all__ = ["GeneratorModel", "DiscriminatorModel", "compile_srgan_model", "train_eval_discriminator",
           "train_eval_generator", "trainer", "GraphedTrainStep", "predict_continent", "HostBand", "HostDEM", "Adam", "DeepBedMapError"]


def __getattr__(name):  # lazy: importing the package must not require a GPU
    if name in ("GeneratorModel", "DiscriminatorModel", "Variable"):
        from . import model
        return getattr(model, name)
    if name in ("compile_srgan_model", "train_eval_discriminator", "train_eval_generator", "trainer", "Adam",
                "ArrayIterator", "DeviceArrayIterator", "save_model_weights_and_architecture", "GraphedTrainStep",
                "set_deterministic"):
        from . import train
        return getattr(train, name)
    if name in ("predict_continent", "tile_plan", "ContinentGrids", "HostBand", "HostDEM"):
        from . import tiler
        return getattr(tiler, name)
    raise AttributeError(name)
