"""PAVRM latent reward scoring (train_pavrm.py:792-845, train_prfl.py:764-796, inference_pavrm.py:575):
truncated Wan-DiT (first `num_blocks` blocks, head removed) -> features after the last kept block ->
QueryAttention pooling -> MLP -> reward logit.  This is the public call bench.py times."""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .model import WanModel
from .network import MLP, QueryAttention


class PavrmScorer(nn.Module):
    def __init__(self, transformer: WanModel, query_attention: QueryAttention, mlp: MLP, feature_layer: int):
        super().__init__()
        self.transformer = transformer
        self.query_attention = query_attention
        self.mlp = mlp
        self.feature_layer = feature_layer

    @classmethod
    def from_state_dicts(cls, cfg: dict, sd_model: Dict[str, torch.Tensor], sd_qa, sd_mlp, num_blocks: int = 8,
                         qa_heads: int = 8, device="cuda"):
        cfg = dict(cfg)
        m = WanModel(**cfg)
        m.load_state_dict(sd_model, strict=True)
        # keep blocks [0, num_blocks) and drop the head, as train_pavrm.py:215-235 does
        m.blocks = nn.ModuleList([m.blocks[i] for i in range(num_blocks)])
        m.head = None
        qa = QueryAttention(cfg["dim"], num_queries=1, num_heads=qa_heads, dropout=0.0, return_type="query")
        qa.load_state_dict(sd_qa, strict=True)
        mlp = MLP(cfg["dim"])
        mlp.load_state_dict(sd_mlp, strict=True)
        return cls(m, qa, mlp, num_blocks).to(device).eval()

    def features(self, x: List[torch.Tensor], t, context, seq_len, clip_fea=None, y=None, gather: bool = True):
        feats = self.transformer(x=x, t=t, context=context, seq_len=seq_len, clip_fea=clip_fea, y=y,
                                 output_features=True, selected_layers=[self.feature_layer], gather_features=gather)
        return torch.stack(feats)                                    # list2batch: [n_sel, B, L, C]

    @torch.no_grad()
    def score(self, x, t, context, seq_len, clip_fea=None, y=None, return_features: bool = False):
        from .parallel import get_sequence_parallel_state
        if get_sequence_parallel_state() and not return_features and os.environ.get("PRFL_SP_POOL", "local") != "gather":
            # sequence parallel: pool each rank's token chunk and merge the partial softmax poolings (no feature all-gather)
            feats = self.features(x, t, context, seq_len, clip_fea, y, gather=False)
            return self.mlp(self.query_attention(feats, sp_local=True))
        feats = self.features(x, t, context, seq_len, clip_fea, y)
        logit = self.mlp(self.query_attention(feats))                # [B, 1, 1]
        return (logit, feats) if return_features else logit
