"""``models.modules`` of the reference -> B200-native ``get_mlp`` / ``LanguageEncoder``."""
from text2pos_cvpr2022_b200.modules import LanguageEncoder, get_mlp  # noqa: F401
