"""``models.pointcloud.pointnet2`` of the reference -> B200-native ``PointNet2``."""
from text2pos_cvpr2022_b200.pointnet2 import GlobalAbstractionLayer, PointNet2, SetAbstractionLayer  # noqa: F401
