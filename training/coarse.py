"""``from training.coarse import eval_epoch`` (reference ``evaluation/pipeline.py:28``) -> the B200-native drop-in."""
from text2pos_cvpr2022_b200.coarse_eval import eval_epoch  # noqa: F401
