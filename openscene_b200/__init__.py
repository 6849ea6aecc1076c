"""openscene_b200: B200-native sparse-3D-convolution + open-vocabulary matching engine behind
OpenScene's MinkUNet (models/mink_unet.py) and run/evaluate.py.  See DESIGN.md."""
__all__ = ['me', 'coords', 'minkunet', 'synth']
