"""``models.superglue_matcher`` of the reference -> B200-native ``SuperGlueMatch`` and ``get_pos_in_cell``."""
from text2pos_cvpr2022_b200.superglue_matcher import SuperGlueMatch, get_mlp_offset, get_pos_in_cell  # noqa: F401
