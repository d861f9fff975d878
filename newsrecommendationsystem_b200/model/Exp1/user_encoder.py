"""Exp1 UserEncoder (reference src/model/Exp1/user_encoder.py:7-31) on libnrms_b200: NRMS's user encoder over
user_vector + position_embedding."""
import torch
import torch.nn as nn

from ... import ops
from ...config import resolve_mode
from ..general.attention.multihead_self import MultiHeadSelfAttention
from ..general.attention.additive import AdditiveAttention


class UserEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.multihead_self_attention = MultiHeadSelfAttention(config.word_embedding_dim, config.num_attention_heads)
        self.position_embedding = nn.Parameter(
            torch.empty(config.num_clicked_news_a_user, config.word_embedding_dim).uniform_(-0.1, 0.1))
        self.additive_attention = AdditiveAttention(config.query_vector_dim, config.word_embedding_dim)
        self.precision = None

    def forward(self, user_vector):
        """user_vector: batch_size, num_clicked_news_a_user, word_embedding_dim -> batch_size, word_embedding_dim"""
        dev = self.position_embedding.device
        x = ops.add_position(user_vector.to(dev), self.position_embedding)
        wqkv, bqkv = self.multihead_self_attention.packed()
        a = self.additive_attention
        return ops.user_encoder(x, wqkv, bqkv, a.linear.weight, a.linear.bias, a.attention_query_vector,
                                mode=resolve_mode(self.config, self.precision))
