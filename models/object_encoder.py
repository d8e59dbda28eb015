"""``models.object_encoder`` of the reference -> B200-native ``ObjectEncoder``."""
from text2pos_cvpr2022_b200.object_encoder import ObjectEncoder  # noqa: F401
