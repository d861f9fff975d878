"""Exp1 NewsEncoder (reference src/model/Exp1/news_encoder.py:10-111) on libnrms_b200.

TextEncoder (:10-34) is NRMS's news encoder (embedding gather, dropout, multi-head self-attention, dropout, additive
attention) with its own attention weights over a SHARED word embedding; ElementEncoder (:37-44) is
relu(linear(embedding(element))) over a shared category embedding; the final AdditiveAttention (:104-110) pools the
2..4 vectors.  Sub-module names follow the reference so that state_dict keys are identical."""
import torch
import torch.nn as nn

from ... import ops
from ...config import resolve_mode
from ..general.attention.multihead_self import MultiHeadSelfAttention
from ..general.attention.additive import AdditiveAttention


class TextEncoder(nn.Module):
    def __init__(self, word_embedding, word_embedding_dim, num_attention_heads, query_vector_dim, dropout_probability):
        super().__init__()
        self.word_embedding = word_embedding
        self.dropout_probability = dropout_probability
        self.multihead_self_attention = MultiHeadSelfAttention(word_embedding_dim, num_attention_heads)
        self.additive_attention = AdditiveAttention(query_vector_dim, word_embedding_dim)
        self.precision = None
        self._dropout_calls = 0
        self.dropout_seed = 0x5EED

    def _seed(self):
        import torch.distributed as dist
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        return (int(self.dropout_seed) + 0x9E3779B97F4A7C15 * rank) & 0xFFFFFFFFFFFFFFFF

    def forward(self, text):
        """text: integer [n, num_words_text] (20 = title; 50 = abstract, inference only) -> fp32 [n, 300]."""
        dev = self.word_embedding.weight.device
        if not text.is_cuda and text.numel():
            lo, hi = int(text.min()), int(text.max())
            if lo < 0 or hi >= self.word_embedding.num_embeddings:
                raise IndexError(f"token ids must lie in [0, {self.word_embedding.num_embeddings}), got [{lo}, {hi}]")
        text = text.to(dev, non_blocking=True)
        wqkv, bqkv = self.multihead_self_attention.packed()
        p = float(self.dropout_probability) if self.training else 0.0
        offset = 0
        if p > 0.0:
            offset = self._dropout_calls
            self._dropout_calls += (text.numel() * ops.D) // 4 + 1
        a = self.additive_attention
        return ops.news_encoder(text, self.word_embedding.weight, wqkv, bqkv, a.linear.weight, a.linear.bias,
                                a.attention_query_vector, dropout_p=p, seed=self._seed(), offset=offset,
                                mode=resolve_mode(None, self.precision))


class ElementEncoder(nn.Module):
    def __init__(self, embedding, linear_input_dim, linear_output_dim):
        super().__init__()
        self.embedding = embedding
        self.linear = nn.Linear(linear_input_dim, linear_output_dim)

    def forward(self, element):
        """element: integer [n] -> fp32 [n, 300]"""
        dev = self.linear.weight.device
        if not element.is_cuda and element.numel():
            lo, hi = int(element.min()), int(element.max())
            if lo < 0 or hi >= self.embedding.num_embeddings:
                raise IndexError(f"element ids must lie in [0, {self.embedding.num_embeddings}), got [{lo}, {hi}]")
        return ops.element_encoder(element.to(dev, non_blocking=True), self.embedding.weight, self.linear.weight,
                                   self.linear.bias)


class NewsEncoder(nn.Module):
    def __init__(self, config, pretrained_word_embedding):
        super().__init__()
        self.config = config
        if (config.word_embedding_dim, config.num_attention_heads, config.query_vector_dim,
                config.category_embedding_dim) != (ops.D, ops.H, ops.QD, 100):
            raise RuntimeError("libnrms_b200 is compiled for word_embedding_dim=300, num_attention_heads=15, "
                               "query_vector_dim=200, category_embedding_dim=100 (reference src/config.py:33-45)")
        if pretrained_word_embedding is None:
            word_embedding = nn.Embedding(config.num_words, config.word_embedding_dim, padding_idx=0)
        else:
            word_embedding = nn.Embedding.from_pretrained(pretrained_word_embedding, freeze=False, padding_idx=0)
        assert len(config.dataset_attributes['news']) > 0
        # sorted (the reference iterates a set intersection, :64-70, whose order is not defined): the order only names
        # the rows of the final attention's input, and additive pooling is invariant to it
        text_names = sorted(set(config.dataset_attributes['news']) & {'title', 'abstract'})
        self.text_encoders = nn.ModuleDict({
            name: TextEncoder(word_embedding, config.word_embedding_dim, config.num_attention_heads,
                              config.query_vector_dim, config.dropout_probability)
            for name in text_names
        })
        category_embedding = nn.Embedding(config.num_categories, config.category_embedding_dim, padding_idx=0)
        element_names = sorted(set(config.dataset_attributes['news']) & {'category', 'subcategory'})
        self.element_encoders = nn.ModuleDict({
            name: ElementEncoder(category_embedding, config.category_embedding_dim, config.word_embedding_dim)
            for name in element_names
        })
        if len(config.dataset_attributes['news']) > 1:
            self.final_attention = AdditiveAttention(config.query_vector_dim, config.word_embedding_dim)
        self._precision = None

    @property
    def precision(self):
        return self._precision

    @precision.setter
    def precision(self, name):
        self._precision = name
        for enc in self.text_encoders.values():
            enc.precision = name
        if hasattr(self, "final_attention"):
            self.final_attention.precision = name

    def attributes(self):
        return list(self.text_encoders.keys()) + list(self.element_encoders.keys())

    def forward(self, news):
        """news: {"category": B, "subcategory": B, "title": B x num_words_title, "abstract": B x num_words_abstract}
        -> B x word_embedding_dim"""
        all_vectors = [encoder(news[name]) for name, encoder in self.text_encoders.items()] + \
                      [encoder(news[name]) for name, encoder in self.element_encoders.items()]
        if len(all_vectors) == 1:
            return all_vectors[0]
        return self.final_attention(ops.stack_vectors(all_vectors))
