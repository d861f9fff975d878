"""DotProductClickPredictor (reference src/model/general/click_predictor/dot_product.py:4-19)."""
import torch

from .... import ops


class DotProductClickPredictor(torch.nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, candidate_news_vector, user_vector):
        """candidate_news_vector [B,C,X], user_vector [B,X] -> [B,C]"""
        return ops.click_score(candidate_news_vector, user_vector)
