"""``models.superglue`` of the reference -> B200-native ``SuperGlue``."""
from text2pos_cvpr2022_b200.superglue import SuperGlue  # noqa: F401
