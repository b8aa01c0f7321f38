"""CPU/PyTorch ORACLE for the Transformer-XL hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference (StefanHeng/Symbolic-Music-Generation) holds no golden
vectors, known-answer tests or fixtures for this path, and its
arithmetic lives in the un-vendored third-party `transformers==4.25.1`
(reference `requirements.txt:150`, module `transformers/models/transfo_xl/
modeling_transfo_xl{,_utilities}.py`) which cannot be imported here (the installed
transformers 5.5 dropped TransfoXL; no wheel, no network).  This file is therefore a
*restatement of the published HF 4.25.1 algorithm* (SURVEY.md Appendix A), anchored on
the reference's own call sites:

  * config derivation ............ musicnlp/models/transformer_xl.py:15-77
  * forward override ............. musicnlp/models/transformer_xl.py:130-221
  * generation inputs ............ musicnlp/models/transformer_xl.py:223-241
  * generate kwargs .............. musicnlp/trainer/eval.py:277-333
  * logged config / param count .. notebook/train/transformer-xl.ipynb:491-580

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / --impl reference
legs may import this module.  The product package never does; it fails loudly when its
CUDA extension is missing.

The implementation deliberately keeps HF's *literal* constructions (time-major tensors,
`torch.cat([mems, w])`, the zero-pad + `view` `_rel_shift`, the uint8 `triu + tril` mask,
`masked_fill` + dense softmax) so that the closed forms used by the CUDA kernels are
checked against the original trick rather than against themselves.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

PT_LOSS_PAD = -100  # musicnlp/util/train/train_util_wrap.py:22


# ----------------------------------------------------------------------------- config
# musicnlp/models/transformer_xl.py:16-49 (presets) + HF TransfoXLConfig defaults that the
# reference leaves untouched (confirmed by the dump at notebook/train/transformer-xl.ipynb:492-513).
_PRESETS = {
    'debug': dict(d_model=128, n_head=8, n_layer=4),
    'debug-large': dict(d_model=128, n_head=8, n_layer=4),
    'tiny': dict(d_model=256, n_head=8, n_layer=6),
    'small': dict(d_model=512, n_head=8, n_layer=12),
    'base': dict(d_model=768, n_head=12, n_layer=12),
    'large': dict(d_model=1024, n_head=16, n_layer=18),
}
_SIZE2MAXLEN = {'debug': 64, 'debug-large': 128, 'tiny': 512, 'small': 1024, 'base': 2048, 'large': 2048}


@dataclass
class RefConfig:
    vocab_size: int = 1190
    d_model: int = 512
    n_head: int = 8
    n_layer: int = 12
    d_head: int = 64
    d_inner: int = 2048
    d_embed: int = 512
    mem_len: int = 1024
    clamp_len: int = 1024
    same_length: bool = True
    dropout: float = 0.1
    dropatt: float = 0.0
    layer_norm_epsilon: float = 1e-5
    init_std: float = 0.02
    eos_token_id: int = 0
    cutoffs: list = field(default_factory=list)
    max_length_: int = 1024

    @classmethod
    def from_preset(cls, model_size: str = 'base', vocab_size: int = 1190, max_length: int = None, **kw):
        """musicnlp/models/transformer_xl.py:26-49,53-70."""
        p = dict(_PRESETS[model_size])
        d, h = p['d_model'], p['n_head']
        assert d % h == 0
        if 'debug' in model_size:
            m_len, c_len = 64, 64
        else:
            m_len = max(128, _SIZE2MAXLEN[model_size] // 8)
            c_len = max(1024, _SIZE2MAXLEN[model_size] // 2)
        p.update(d_embed=d, d_inner=d * 4, d_head=d // h, mem_len=m_len, clamp_len=c_len, vocab_size=vocab_size)
        p['cutoffs'] = [1000] if vocab_size >= 1000 else []   # :56-66 (small-vocab rows only)
        p.update(kw)
        p['max_length_'] = max_length or _SIZE2MAXLEN[model_size]
        return cls(**p)


# ----------------------------------------------------------------------------- literal index constructions
def literal_rel_shift(x: torch.Tensor) -> torch.Tensor:
    """HF `_rel_shift` (Appendix A.4): x is (qlen, klen, ...)."""
    zero_pad_shape = (x.size(0), 1) + x.size()[2:]
    zero_pad = torch.zeros(zero_pad_shape, device=x.device, dtype=x.dtype)
    x_padded = torch.cat([zero_pad, x], dim=1)
    x_padded_shape = (x.size(1) + 1, x.size(0)) + x.size()[2:]
    x_padded = x_padded.view(*x_padded_shape)
    return x_padded[1:].view_as(x)


def literal_attn_mask(qlen: int, mlen: int, mem_len: int, same_length: bool) -> torch.Tensor:
    """HF `TransfoXLModel.forward` mask (Appendix A.2-4): uint8 (qlen, klen), 1 = masked."""
    klen = mlen + qlen
    all_ones = torch.ones((qlen, klen), dtype=torch.uint8)
    if same_length:
        mask_len = klen - mem_len
        mask_shift_len = qlen - mask_len if mask_len > 0 else qlen
        return torch.triu(all_ones, 1 + mlen) + torch.tril(all_ones, -mask_shift_len)
    return torch.triu(all_ones, diagonal=1 + mlen)


def literal_pos_seq(klen: int, clamp_len: int) -> torch.Tensor:
    """`pos_seq = arange(klen-1, -1, -1.0)`, clamped (Appendix A.2-5).  Returned as int64."""
    pos_seq = torch.arange(klen - 1, -1, -1, dtype=torch.int64)
    if clamp_len > 0:
        pos_seq = pos_seq.clamp(max=clamp_len)
    return pos_seq


def literal_index_maps(qlen: int, mlen: int, mem_len: int, clamp_len: int, same_length: bool = True):
    """Integer ground truth for the CUDA kernels' closed forms.

    Returns (masked uint8 (qlen,klen), ridx int64 (qlen,klen)) where `ridx[i,j]` is the *relative
    position value* whose embedding is multiplied with query i at key j after the literal
    pad/reshape shift (-1 where the shifted element is the zero pad, -2 - p where it wrapped from the
    next query row; both only ever occur at masked places).
    """
    klen = mlen + qlen
    pos = literal_pos_seq(klen, clamp_len)                      # (klen,)
    # code every BD0[i, j'] by (i, pos[j']) -> shift -> decode
    code = (torch.arange(qlen).view(-1, 1) * (klen + 7) + pos.view(1, -1) + 1).to(torch.int64)   # >0
    shifted = literal_rel_shift(code)
    row = (shifted - 1) // (klen + 7)
    p = (shifted - 1) % (klen + 7)
    own = row == torch.arange(qlen).view(-1, 1)
    ridx = torch.where(shifted == 0, torch.full_like(p, -1), torch.where(own, p, -2 - p))
    return literal_attn_mask(qlen, mlen, mem_len, same_length), ridx


# ----------------------------------------------------------------------------- modules (HF state_dict names, Appendix A.8)
class _PosEmb(nn.Module):
    def __init__(self, demb):
        super().__init__()
        inv_freq = 1 / (10000 ** (torch.arange(0.0, demb, 2.0) / demb))
        self.register_buffer('inv_freq', inv_freq)

    def forward(self, pos_seq):
        sinusoid = torch.outer(pos_seq, self.inv_freq)
        return torch.cat([sinusoid.sin(), sinusoid.cos()], dim=-1)[:, None, :]


class _FF(nn.Module):
    def __init__(self, d, di, p, eps):
        super().__init__()
        self.CoreNet = nn.Sequential(nn.Linear(d, di), nn.ReLU(inplace=True), nn.Dropout(p),
                                     nn.Linear(di, d), nn.Dropout(p))
        self.layer_norm = nn.LayerNorm(d, eps=eps)

    def forward(self, inp):
        return self.layer_norm(inp + self.CoreNet(inp))


class _RelAttn(nn.Module):
    def __init__(self, cfg: RefConfig):
        super().__init__()
        H, dh, d = cfg.n_head, cfg.d_head, cfg.d_model
        self.n_head, self.d_head, self.scale = H, dh, 1 / (dh ** 0.5)
        self.qkv_net = nn.Linear(d, 3 * H * dh, bias=False)
        self.r_net = nn.Linear(d, H * dh, bias=False)
        self.o_net = nn.Linear(H * dh, d, bias=False)
        self.drop = nn.Dropout(cfg.dropout)
        self.dropatt = nn.Dropout(cfg.dropatt)
        self.layer_norm = nn.LayerNorm(d, eps=cfg.layer_norm_epsilon)
        self.r_r_bias = nn.Parameter(torch.zeros(H, dh))   # untie_r=True -> per layer
        self.r_w_bias = nn.Parameter(torch.zeros(H, dh))

    def forward(self, w, r, attn_mask, mems, return_prob=False):
        qlen, rlen, bsz = w.size(0), r.size(0), w.size(1)
        H, dh = self.n_head, self.d_head
        cat = torch.cat([mems, w], 0) if mems is not None else w
        w_heads = self.qkv_net(cat)
        r_head_k = self.r_net(r)
        w_head_q, w_head_k, w_head_v = torch.chunk(w_heads, 3, dim=-1)
        w_head_q = w_head_q[-qlen:]
        klen = w_head_k.size(0)
        w_head_q = w_head_q.view(qlen, bsz, H, dh)
        w_head_k = w_head_k.view(klen, bsz, H, dh)
        w_head_v = w_head_v.view(klen, bsz, H, dh)
        r_head_k = r_head_k.view(rlen, H, dh)
        AC = torch.einsum('ibnd,jbnd->ijbn', w_head_q + self.r_w_bias, w_head_k)
        BD = torch.einsum('ibnd,jnd->ijbn', w_head_q + self.r_r_bias, r_head_k)
        BD = literal_rel_shift(BD)
        attn_score = (AC + BD) * self.scale
        if attn_mask is not None and bool(attn_mask.any()):
            neg = torch.finfo(attn_score.dtype).min
            attn_score = attn_score.float().masked_fill(attn_mask[:, :, None, None].bool(), neg).type_as(attn_score)
        attn_prob = self.dropatt(F.softmax(attn_score, dim=1))
        attn_vec = torch.einsum('ijbn,jbnd->ibnd', attn_prob, w_head_v)
        attn_vec = attn_vec.contiguous().view(qlen, bsz, H * dh)
        attn_out = self.drop(self.o_net(attn_vec))
        out = self.layer_norm(w + attn_out)
        return (out, attn_prob) if return_prob else out


class _Layer(nn.Module):
    def __init__(self, cfg: RefConfig):
        super().__init__()
        self.dec_attn = _RelAttn(cfg)
        self.pos_ff = _FF(cfg.d_model, cfg.d_inner, cfg.dropout, cfg.layer_norm_epsilon)

    def forward(self, x, r, mask, mems):
        return self.pos_ff(self.dec_attn(x, r, mask, mems))


class _AdaptiveEmb(nn.Module):
    def __init__(self, V, d):
        super().__init__()
        self.emb_scale = d ** 0.5
        self.emb_layers = nn.ModuleList([nn.Embedding(V, d)])

    def forward(self, ids):
        return self.emb_layers[0](ids) * self.emb_scale


class _Crit(nn.Module):
    """HF 4.25.1 `ProjectedAdaptiveLogSoftmax` at div_val == 1, d_proj == d_embed (Appendix A.6): the n_clusters == 0 path the reference
    trains with (`cutoffs=[]`, musicnlp/trainer/train.py:521-527) AND the cluster path its config derivation selects by default for
    vocabularies >= 1000 (`cutoffs=[1000]` ..., musicnlp/models/transformer_xl.py:56-66), restated construct by construct:
    `cutoffs + [n_token]`, head = shortlist rows + `cluster_weight`, `cluster_prob_idx = cutoffs[0] + i - 1`, per-cluster
    `index_select`, and - because TransfoXLLMHeadModel builds the criterion with keep_order=False - the PACKED loss vector
    (`out[offset : offset + n_i]`, clusters in order, ignored labels left as trailing zeros)."""

    def __init__(self, V, d, cutoffs=(), keep_order=False):
        super().__init__()
        self.n_token = V
        self.cutoffs = list(cutoffs) + [V]
        self.cutoff_ends = [0] + self.cutoffs
        self.shortlist_size = self.cutoffs[0]
        self.n_clusters = len(self.cutoffs) - 1
        self.head_size = self.shortlist_size + self.n_clusters
        if self.n_clusters > 0:
            self.cluster_weight = nn.Parameter(torch.zeros(self.n_clusters, d))
            self.cluster_bias = nn.Parameter(torch.zeros(self.n_clusters))
        self.out_layers = nn.ModuleList([nn.Linear(d, V)])
        self.keep_order = keep_order

    def forward(self, hidden, labels=None, keep_order=False):
        if labels is not None:
            hidden = hidden[..., :-1, :].contiguous()
            labels = labels[..., 1:].contiguous()
            hidden = hidden.view(-1, hidden.size(-1))
            labels = labels.view(-1)
            if hidden.size(0) != labels.size(0):
                raise RuntimeError('Input and labels should have the same size in the batch dimension.')
        else:
            hidden = hidden.view(-1, hidden.size(-1))
        if self.n_clusters == 0:
            logit = F.linear(hidden, self.out_layers[0].weight, self.out_layers[0].bias)
            if labels is not None:
                mask = labels != PT_LOSS_PAD
                out = torch.zeros_like(labels, dtype=hidden.dtype)
                out[mask] = -F.log_softmax(logit, dim=-1)[mask].gather(1, labels[mask].unsqueeze(1)).squeeze(1)
                return out
            return F.log_softmax(logit, dim=-1)
        # ---- cluster path
        weights, biases = [], []
        for i in range(len(self.cutoffs)):
            l_idx, r_idx = self.cutoff_ends[i], self.cutoff_ends[i + 1]
            weight_i = self.out_layers[0].weight[l_idx:r_idx]
            bias_i = self.out_layers[0].bias[l_idx:r_idx]
            if i == 0:
                weight_i = torch.cat([weight_i, self.cluster_weight], dim=0)
                bias_i = torch.cat([bias_i, self.cluster_bias], dim=0)
            weights.append(weight_i)
            biases.append(bias_i)
        head_logit = F.linear(hidden, weights[0], biases[0])
        head_logprob = F.log_softmax(head_logit, dim=1)
        if labels is None:
            out = hidden.new_empty((head_logit.size(0), self.n_token))
        else:
            out = torch.zeros_like(labels, dtype=hidden.dtype)
        offset = 0
        cutoff_values = [0] + self.cutoffs
        for i in range(len(cutoff_values) - 1):
            l_idx, r_idx = cutoff_values[i], cutoff_values[i + 1]
            if labels is not None:
                mask_i = (labels >= l_idx) & (labels < r_idx)
                indices_i = mask_i.nonzero().squeeze()
                if indices_i.numel() == 0:
                    continue
                indices_i = indices_i.view(-1)          # (.squeeze() of a single hit is 0-d; index_select wants 1-d - same values)
                target_i = labels.index_select(0, indices_i) - l_idx
                head_logprob_i = head_logprob.index_select(0, indices_i)
                hidden_i = hidden.index_select(0, indices_i)
            else:
                hidden_i = hidden
            if i == 0:
                if labels is not None:
                    logprob_i = head_logprob_i.gather(1, target_i[:, None]).squeeze(1)
                else:
                    out[:, :self.cutoffs[0]] = head_logprob[:, :self.cutoffs[0]]
            else:
                tail_logit_i = F.linear(hidden_i, weights[i], biases[i])
                tail_logprob_i = F.log_softmax(tail_logit_i, dim=1)
                cluster_prob_idx = self.cutoffs[0] + i - 1          # no probability for the head cluster
                if labels is not None:
                    logprob_i = head_logprob_i[:, cluster_prob_idx] + tail_logprob_i.gather(1, target_i[:, None]).squeeze(1)
                else:
                    logprob_i = head_logprob[:, cluster_prob_idx, None] + tail_logprob_i
                    out[:, l_idx:r_idx] = logprob_i
            if labels is not None:
                if self.keep_order or keep_order:
                    out.index_copy_(0, indices_i, -logprob_i)
                else:
                    out[offset:offset + logprob_i.size(0)].copy_(-logprob_i)
                offset += logprob_i.size(0)
        return out


class _Transformer(nn.Module):
    def __init__(self, cfg: RefConfig):
        super().__init__()
        self.cfg = cfg
        self.word_emb = _AdaptiveEmb(cfg.vocab_size, cfg.d_model)
        self.drop = nn.Dropout(cfg.dropout)
        self.layers = nn.ModuleList([_Layer(cfg) for _ in range(cfg.n_layer)])
        self.pos_emb = _PosEmb(cfg.d_model)

    def init_mems(self, bsz):
        if self.cfg.mem_len > 0:
            p = next(self.parameters())
            return [torch.zeros(self.cfg.mem_len, bsz, self.cfg.d_model, dtype=p.dtype, device=p.device)
                    for _ in range(self.cfg.n_layer)]
        return None

    def _update_mems(self, hids, mems, mlen, qlen):
        if mems is None:
            return None
        with torch.no_grad():
            end_idx = mlen + max(0, qlen)
            beg_idx = max(0, end_idx - self.cfg.mem_len)
            return [torch.cat([mems[i], hids[i]], dim=0)[beg_idx:end_idx].detach() for i in range(len(hids))]

    def forward(self, input_ids, mems=None):
        cfg = self.cfg
        ids = input_ids.transpose(0, 1).contiguous()
        qlen, bsz = ids.size()
        if mems is None:
            mems = self.init_mems(bsz)
        word_emb = self.word_emb(ids)
        mlen = mems[0].size(0) if mems is not None else 0
        klen = mlen + qlen
        mask = literal_attn_mask(qlen, mlen, cfg.mem_len, cfg.same_length).to(word_emb.device)
        pos_seq = torch.arange(klen - 1, -1, -1.0, device=word_emb.device, dtype=word_emb.dtype)
        if cfg.clamp_len > 0:
            pos_seq.clamp_(max=cfg.clamp_len)
        pos_emb = self.pos_emb(pos_seq)
        core_out = self.drop(word_emb)
        pos_emb = self.drop(pos_emb)
        hids = []
        for i, layer in enumerate(self.layers):
            hids.append(core_out)
            core_out = layer(core_out, pos_emb, mask, None if mems is None else mems[i])
        core_out = self.drop(core_out)
        new_mems = self._update_mems(hids, mems, mlen, qlen)
        return core_out.transpose(0, 1).contiguous(), new_mems


class RefOutput(dict):
    """reference musicnlp/models/transformer_xl.py:81-124 (`TransfoXLLMHeadModelOutput`, an HF `ModelOutput`): dict-like over the non-None
    fields in the order `losses, prediction_scores, mems, hidden_states, attentions, loss`, with attribute access and `.logits`."""
    _fields = ('losses', 'prediction_scores', 'mems', 'hidden_states', 'attentions', 'loss')

    def __init__(self, loss=None, losses=None, prediction_scores=None, mems=None, hidden_states=None, attentions=None):
        super().__init__()
        vals = dict(losses=losses, prediction_scores=prediction_scores, mems=mems, hidden_states=hidden_states, attentions=attentions, loss=loss)
        for k in self._fields:
            object.__setattr__(self, k, vals[k])
            if vals[k] is not None:
                self[k] = vals[k]

    @property
    def logits(self):
        return self.prediction_scores


class RefTransfoXLLMHeadModel(nn.Module):
    """Restated `MyTransfoXLLMHeadModel` (reference musicnlp/models/transformer_xl.py:127-241)."""

    def __init__(self, cfg: RefConfig):
        super().__init__()
        self.config = cfg
        self.transformer = _Transformer(cfg)
        self.crit = _Crit(cfg.vocab_size, cfg.d_model, cfg.cutoffs)      # HF builds it with keep_order=False
        self.apply(self._init_weights)
        self.crit.out_layers[0].weight = self.transformer.word_emb.emb_layers[0].weight   # tie_word_embeddings

    def _init_weights(self, m):
        """Appendix A.9 (`init='normal'`, std 0.02)."""
        std = self.config.init_std
        if isinstance(m, nn.Linear):
            nn.init.normal_(m.weight, 0.0, std)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0.0)
        elif isinstance(m, nn.Embedding):
            nn.init.normal_(m.weight, 0.0, std)
        elif isinstance(m, nn.LayerNorm):
            nn.init.normal_(m.weight, 1.0, std)
            nn.init.constant_(m.bias, 0.0)
        elif isinstance(m, _RelAttn):
            nn.init.normal_(m.r_w_bias, 0.0, std)
            nn.init.normal_(m.r_r_bias, 0.0, std)
        elif isinstance(m, _Crit) and m.n_clusters > 0:
            nn.init.normal_(m.cluster_weight, 0.0, std)
            nn.init.constant_(m.cluster_bias, 0.0)

    def num_parameters(self):
        return sum(p.numel() for p in self.parameters())    # tied weight counted once by .parameters()

    def forward(self, input_ids=None, mems=None, labels=None, key_scores=None) -> RefOutput:
        """reference transformer_xl.py:130-221."""
        if input_ids is None:
            raise ValueError('You have to specify either input_ids or inputs_embeds')
        bsz, tgt_len = input_ids.size(0), input_ids.size(1)
        last_hidden, new_mems = self.transformer(input_ids, mems=mems)
        pred_hid = last_hidden[:, -tgt_len:]
        if labels is not None:
            miss_valid_label = labels[0, 1:].sum() == (labels.size(1) - 1) * -100      # :176-182
            if miss_valid_label:
                labels[0, 1] = self.config.eos_token_id
        softmax_output = self.crit(pred_hid, labels)
        _softmax_output = softmax_output
        in_eval = not self.training
        if in_eval and labels is not None:
            _softmax_output = self.crit(pred_hid, None)
        prediction_scores = _softmax_output.view(bsz, tgt_len, -1) if (labels is None or in_eval) else ()
        if labels is not None:
            losses = softmax_output.view(bsz, tgt_len - 1)
            loss = losses[losses != 0].mean()
        else:
            losses, loss = None, None
        return RefOutput(loss=loss, losses=losses, prediction_scores=prediction_scores, mems=new_mems)

    # ------------------------------------------------------------------ generation (Appendix A.7)
    @staticmethod
    def prepare_inputs_for_generation(input_ids, past=None):
        """reference transformer_xl.py:223-241."""
        if past:
            return dict(mems=past, input_ids=input_ids[:, -1].unsqueeze(-1))
        return dict(input_ids=input_ids)

    @staticmethod
    def warp_scores(scores, temperature=1.0, top_k=0, top_p=1.0, renormalize=True):
        """Temperature -> TopK -> TopP -> log_softmax, HF 4.25 order; returns warped log-probs (B,V)."""
        s = scores.clone()
        neg_inf = -float('inf')
        if temperature != 1.0:
            s = s / temperature
        if top_k and top_k > 0:
            k = min(max(top_k, 1), s.size(-1))
            kth = torch.topk(s, k)[0][..., -1, None]
            s = s.masked_fill(s < kth, neg_inf)
        if top_p is not None and top_p < 1.0:
            sorted_logits, sorted_idx = torch.sort(s, descending=True)
            cum = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
            remove = cum > top_p
            remove[..., 1:] = remove[..., :-1].clone()
            remove[..., 0] = False
            remove = remove.scatter(1, sorted_idx, remove)
            s = s.masked_fill(remove, neg_inf)
        if renormalize:
            s = F.log_softmax(s, dim=-1)
        return s

    @staticmethod
    def repetition_penalty(scores, prev_ids, penalty):
        """HF 4.25 RepetitionPenaltyLogitsProcessor (a logits PROCESSOR: runs on the raw step scores, before every warper, in greedy
        search too): the score of every token already in the sequence is divided by `penalty` if positive, multiplied if negative.
        Row-by-row restatement (accepted key of the reference's `sample` strategy, musicnlp/trainer/eval.py:279)."""
        out = scores.clone()
        for b in range(scores.size(0)):
            for t in set(prev_ids[b].tolist()):
                v = scores[b, t]
                out[b, t] = v * penalty if v < 0 else v / penalty
        return out

    @staticmethod
    def typical_filter(scores, mass, min_tokens_to_keep=1):
        """HF 4.25 TypicalLogitsWarper (locally typical sampling; `typical_p`, eval.py:279): keep the tokens whose surprisal is closest to
        the entropy of the distribution until their mass reaches `mass`; the rest -> -inf.  Row-by-row restatement."""
        out = scores.clone()
        for b in range(scores.size(0)):
            logp = F.log_softmax(scores[b].double(), dim=-1)
            pr = logp.exp()
            ent = -torch.nansum(logp * pr)
            dist = ((-logp) - ent).abs().float()          # HF sorts the fp32 distances
            order = torch.sort(dist, stable=False)[1]
            cum = torch.softmax(scores[b][order], dim=-1).cumsum(-1)
            last = int((cum < mass).sum().item())
            last = min(last, scores.size(1) - 1)
            thr = dist[order[last]]
            remove = dist > thr
            if min_tokens_to_keep > 1:
                remove[order[:min_tokens_to_keep]] = False
            out[b, remove] = -float('inf')
        return out

    @torch.no_grad()
    def generate(self, input_ids, max_length, do_sample=False, temperature=1.0, top_k=50, top_p=1.0,
                 renormalize_logits=True, generator=None, eos_token_id=None, pad_token_id=None,
                 return_step_scores=False, typical_p=None, repetition_penalty=None):
        """greedy_search / sample loop of HF GenerationMixin 4.25 restricted to what eval.py:277-333 uses."""
        self.eval()
        ids = input_ids.clone()
        past = None
        unfinished = torch.ones(ids.size(0), dtype=torch.long)
        step_scores = []
        while ids.size(1) < max_length:
            inp = self.prepare_inputs_for_generation(ids, past=past)
            out = self.forward(**inp)
            s = out.logits[:, -1, :]
            past = out.mems
            if repetition_penalty is not None and repetition_penalty != 1.0:
                s = self.repetition_penalty(s, ids, repetition_penalty)
                if not do_sample and renormalize_logits:
                    s = F.log_softmax(s, dim=-1)
            if do_sample:
                typ = typical_p is not None and typical_p < 1.0
                s = self.warp_scores(s, temperature, top_k, top_p, renormalize_logits and not typ)
                if typ:      # HF order: temperature, top-k, top-p, typical, then the renormalisation
                    s = self.typical_filter(s, typical_p)
                    if renormalize_logits:
                        s = F.log_softmax(s, dim=-1)
                probs = F.softmax(s, dim=-1)
                nxt = torch.multinomial(probs, 1, generator=generator).squeeze(1)
            else:
                nxt = torch.argmax(s, dim=-1)
            if return_step_scores:
                step_scores.append(s)
            if eos_token_id is not None:
                pad = eos_token_id if pad_token_id is None else pad_token_id
                nxt = nxt * unfinished + pad * (1 - unfinished)
                unfinished = unfinished.mul((nxt != eos_token_id).long())
            ids = torch.cat([ids, nxt[:, None]], dim=-1)
            if eos_token_id is not None and unfinished.max() == 0:
                break
        return (ids, step_scores) if return_step_scores else ids


def expected_param_count(L, d, di, V, n_clusters=0):
    """Appendix A.8 closed form; KAT: (12, 768, 3072, 418) -> 92,435,362 (log says 92.4M)."""
    return L * (3 * d * d + 2 * d * d + 2 * d + 2 * d + d * di + di + di * d + d + 2 * d) + V * d + V + n_clusters * (d + 1)


def ntp_acc_counts(preds, labels, pad=-100):
    """(matches, non-pad count) of next-token prediction: `preds[:, :-1]` vs `labels[:, 1:]` over `labels != -100`.
    Follows reference musicnlp/util/train/train_util_wrap.py:113-120 (training) and musicnlp/trainer/train.py:279-284 (eval,
    `clm_pred_shifted=False`): ntp_acc = matches / count."""
    p, l = preds[:, :-1], labels[:, 1:]
    msk = l != pad
    return int((p[msk] == l[msk]).sum().item()), int(msk.sum().item())


def clm_collate(input_ids, pad_token_id, pad=-100):
    """HF DataCollatorForLanguageModeling(mlm=False).torch_call on already padded examples, as instantiated at reference
    musicnlp/trainer/train.py:360 (examples padded by the tokenizer, musicnlp/preprocess/dataset.py:361):
    labels = input_ids.clone(); labels[labels == pad_token_id] = -100."""
    labels = input_ids.clone()
    labels[labels == pad_token_id] = pad
    return {'input_ids': input_ids, 'labels': labels}


def truncate_last_bar(ids, sob_token_id):
    """reference musicnlp/trainer/eval.py:178-185 (`MusicGenerator._truncate_last_bar`) for one 1-D id tensor."""
    assert ids.dim() == 1
    idxs = torch.nonzero(ids.eq(sob_token_id)).flatten().tolist()
    assert len(idxs) > 0, 'No start of bar token found when truncate_to_sob enabled'
    return ids[:idxs[-1]].tolist()


def hf_param_groups(model, weight_decay):
    """HF Trainer.create_optimizer: decay every parameter that is not inside an nn.LayerNorm and whose name does not contain "bias"
    (so r_w_bias / r_r_bias / CoreNet biases / crit bias are not decayed); tied weights appear once."""
    ln_params = {id(p) for m in model.modules() if isinstance(m, torch.nn.LayerNorm) for p in m.parameters()}
    seen, decay, no_decay = set(), [], []
    for n, p in model.named_parameters():
        if id(p) in seen:
            continue
        seen.add(id(p))
        (no_decay if (id(p) in ln_params or 'bias' in n) else decay).append(p)
    return [dict(params=decay, weight_decay=weight_decay), dict(params=no_decay, weight_decay=0.0)]


def train_steps(model, batches, total_steps, learning_rate=3e-4, weight_decay=1e-2, warmup_ratio=0.1, max_grad_norm=1.0):
    """The optimisation steps HF Trainer runs for the reference (musicnlp/trainer/train.py:166-190, :79-110; compute_loss + ntp_acc of
    musicnlp/util/train/train_util_wrap.py:88-144): AdamW(torch) + clip_grad_norm_ + cosine schedule with ceil(ratio * steps) warm-up steps,
    scheduler stepped after the optimizer.  Returns one dict per step: loss, learning_rate used, grad_norm before clipping, ntp_acc."""
    import math
    opt = torch.optim.AdamW(hf_param_groups(model, weight_decay), lr=learning_rate, betas=(0.9, 0.999), eps=1e-8)
    warm = math.ceil(total_steps * warmup_ratio)

    def lam(step):
        if step < warm:
            return step / max(1, warm)
        prog = (step - warm) / max(1, total_steps - warm)
        return max(0.0, 0.5 * (1.0 + math.cos(math.pi * prog)))
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lam)
    model.train()
    out = []
    for ids, labels in batches:
        o = model(input_ids=ids, labels=labels.clone())
        with torch.no_grad():
            model.eval()
            preds = model(input_ids=ids).logits.argmax(-1)      # what outputs.logits.argmax(-1) is once logits are returned in training
            model.train()
        hit, cnt = ntp_acc_counts(preds, labels)
        opt.zero_grad()
        o.loss.backward()
        gn = torch.nn.utils.clip_grad_norm_(model.parameters(), max_grad_norm)
        lr_used = opt.param_groups[0]['lr']
        opt.step()
        sched.step()
        out.append(dict(loss=float(o.loss.detach()), learning_rate=lr_used, grad_norm=float(gn), ntp_acc=hit / cnt if cnt else float('nan')))
    return out
