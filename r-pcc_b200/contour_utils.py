"""Mirror of utils/contour_utils.py:178-230 (ContourExtractor); the flood-fill coders there are
unused by the tools and not reproduced."""
from .plugin import contour_utils_cpp


class ContourExtractor:
    @staticmethod
    def extract_contour(idx_map):
        return contour_utils_cpp.extract_contour(idx_map)

    @staticmethod
    def recover_map(contour_map, idx_sequence):
        return contour_utils_cpp.recover_map(contour_map, idx_sequence)
