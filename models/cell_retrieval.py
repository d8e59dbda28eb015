"""``models.cell_retrieval`` of the reference -> B200-native ``CellRetrievalNetwork``."""
from text2pos_cvpr2022_b200.cell_retrieval import CellRetrievalNetwork  # noqa: F401
