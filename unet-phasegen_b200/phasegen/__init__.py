"""phasegen -- host side of the B200-native magnitude -> phase -> waveform path.

``ops``       tensor-level wrappers over the C ABI of libphasegen.so (include/phasegen.h)
``unet``      executor of the U-Net forward on those kernels
``pipeline``  wave -> STFT -> U-Net -> ISTFT -> wave, everything resident on the GPU
``synth``     seeded synthetic clips / weights of the shapes BASELINE.json names
"""
from . import _lib  # noqa: F401
