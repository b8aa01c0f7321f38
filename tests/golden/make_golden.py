"""Regenerates tests/golden/oracle_tiny.pt from the oracle (oracle/txl_ref.py).

The reference's own implementation of this path (HF transformers==4.25.1 TransfoXL) cannot be imported in this
environment, and the reference has no golden vectors, so these fixtures pin the ORACLE against
regressions of itself, not against the reference: parity stays "unpinned" (see oracle/txl_ref.py header).
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.txl_ref import RefConfig, RefTransfoXLLMHeadModel, literal_index_maps  # noqa: E402


def main():
    torch.manual_seed(77)   # musicnlp/util/config.json:135 random-seed
    cfg = dict(vocab_size=61, d_model=32, n_head=4, n_layer=2, d_head=8, d_inner=64, d_embed=32, mem_len=6, clamp_len=4, dropout=0.0)
    m = RefTransfoXLLMHeadModel(RefConfig(**cfg)).eval()
    g = torch.Generator().manual_seed(77)
    ids = torch.randint(0, 61, (3, 9), generator=g)
    labels = ids.clone()
    labels[1, 5:] = -100
    with torch.no_grad():
        o1 = m(input_ids=ids, labels=labels.clone())
        o2 = m(input_ids=ids[:, :4], mems=o1.mems, labels=None)
    gen = m.generate(ids[:, :3], max_length=14)
    maps = {f'{T},{M},{ML},{C}': literal_index_maps(T, M, ML, C) for (T, M, ML, C) in [(7, 5, 5, 3), (1, 8, 8, 4), (5, 0, 4, 2), (4, 2, 6, 3)]}
    torch.save(dict(cfg=cfg, state_dict=m.state_dict(), ids=ids, labels=labels, loss=o1.loss, losses=o1.losses, logits=o1.logits,
                    mems0=o1.mems[0], logits2=o2.logits, greedy=gen, maps=maps),
               os.path.join(os.path.dirname(os.path.abspath(__file__)), 'oracle_tiny.pt'))
    print('loss', o1.loss.item())


if __name__ == '__main__':
    main()
