"""NRMS (reference src/model/NRMS/__init__.py:7-84): same constructor, methods, sub-module names
and state_dict keys; every arithmetic step runs in libnrms_b200 (sm_100a)."""
import torch

from .news_encoder import NewsEncoder
from .user_encoder import UserEncoder
from ..general.click_predictor.dot_product import DotProductClickPredictor


class NRMS(torch.nn.Module):
    """Input 1 + K candidate news and a list of user clicked news, produce the click probability."""

    def __init__(self, config, pretrained_word_embedding=None):
        super().__init__()
        self.config = config
        self.news_encoder = NewsEncoder(config, pretrained_word_embedding)
        self.user_encoder = UserEncoder(config)
        self.click_predictor = DotProductClickPredictor()

    def set_precision(self, name):
        """"tf32" (tcgen05) or "fp32" (CUDA cores, reference-exact)."""
        self.news_encoder.precision = name
        self.user_encoder.precision = name
        return self

    def forward(self, candidate_news, clicked_news):
        """candidate_news: [{"title": B x L}] * (1+K); clicked_news: [{"title": B x L}] * N  ->  B x (1+K).

        The reference runs the news encoder 1+K+N times (list comprehension + torch.stack(dim=1),
        :38-42); the encoder is position-wise, so the titles are stacked FIRST and encoded in one
        launch sequence -- same function, one H2D copy instead of 1+K+N."""
        n_cand = len(candidate_news)
        titles = torch.stack([x["title"] for x in candidate_news] + [x["title"] for x in clicked_news], dim=1)
        return self.forward_tokens(titles, n_cand)

    def forward_tokens(self, titles, n_cand):
        """titles: integer [B, 1+K+N, L] (candidates first) -> logits [B, 1+K]."""
        B, T, L = titles.shape
        vec = self.news_encoder.encode_tokens(titles.reshape(B * T, L)).view(B, T, -1)
        candidate_news_vector = vec[:, :n_cand]
        clicked_news_vector = vec[:, n_cand:]
        user_vector = self.user_encoder(clicked_news_vector)
        return self.click_predictor(candidate_news_vector, user_vector)

    def forward_rows(self, token_table, cand_rows, hist_rows):
        """Index-only minibatch: token_table int64 [N_news, L] on the device, cand_rows [B, 1+K] and
        hist_rows [B, N] news-row indices -> logits [B, 1+K]."""
        B, n_cand = cand_rows.shape
        dev = token_table.device
        rows = torch.cat([cand_rows.to(dev, non_blocking=True), hist_rows.to(dev, non_blocking=True)], dim=1)
        vec = self.news_encoder.encode_tokens(token_table, news_rows=rows.reshape(-1)).view(B, rows.shape[1], -1)
        user_vector = self.user_encoder(vec[:, n_cand:])
        return self.click_predictor(vec[:, :n_cand], user_vector)

    def get_news_vector(self, news):
        """news: {"title": B x L} -> B x word_embedding_dim"""
        return self.news_encoder(news)

    def get_user_vector(self, clicked_news_vector):
        """clicked_news_vector: B x N x word_embedding_dim -> B x word_embedding_dim"""
        return self.user_encoder(clicked_news_vector)

    def get_prediction(self, news_vector, user_vector):
        """news_vector: C x X, user_vector: X -> C"""
        return self.click_predictor(news_vector.unsqueeze(dim=0), user_vector.unsqueeze(dim=0)).squeeze(dim=0)
