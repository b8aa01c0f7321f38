"""the model under the reference's HF-Trainer glue.

`transformers.Trainer` itself cannot be instantiated in this image (transformers 5.5 requires `accelerate>=1.1.0` for TrainingArguments;
not installed, no network) — `test_hf_trainer_unavailable_reason` records that.  What is exercised instead is a literal restatement of the
hooks the reference overrides, around the B200 model as a plain `nn.Module` driven by STOCK torch machinery (autograd `.grad`s,
`torch.optim.AdamW`, `clip_grad_norm_`, `LambdaLR`), i.e. nothing of this repo's trainer / fused optimiser:

  * `MyTrainer.compute_loss` (musicnlp/util/train/train_util_wrap.py:88-144): `outputs = model(**inputs)` with `key_scores` among the inputs,
    `outputs["loss"]`, next-token accuracy of `logits.argmax(-1)` vs the shifted, pad-masked labels;
  * HF `Trainer.training_step` / optimizer step as the reference configures it (musicnlp/trainer/train.py:166-190);
  * `MyEvalTrainer.prediction_step` (musicnlp/util/train/trainer_eval_wrap.py:146-152) with `ignore_keys_for_eval=['losses', 'mems',
    'hidden_states', 'attentions']` (train.py:588), `preprocess_logits_for_metrics = max_out_logits` (train.py:248) and the
    `ComputeMetrics` ntp_acc arithmetic (train.py:279-284);
  * `save_pretrained` -> `from_pretrained` (train.py:591, eval.py:91) in the middle of it.
Every logged quantity is compared with the same loop around the oracle."""
import math

import pytest
import torch

from conftest import make_pair
from oracle.txl_ref import hf_param_groups

PT_LOSS_PAD = -100
IGNORE_KEYS_FOR_EVAL = ['losses', 'mems', 'hidden_states', 'attentions']      # musicnlp/trainer/train.py:588


def test_hf_trainer_unavailable_reason():
    """Not gpu-marked: documents why the drive-through below restates the hooks instead of instantiating transformers.Trainer."""
    import importlib.util
    if importlib.util.find_spec('accelerate') is not None:
        pytest.skip('accelerate is installed here: transformers.Trainer could be used directly')
    from transformers import TrainingArguments
    with pytest.raises(ImportError):
        TrainingArguments(output_dir='/tmp/_txl_hf_args', report_to=[])


class _MyTrainerLike:
    """compute_loss / training_step / prediction_step of the reference's trainers, restated (no callbacks, no logging sinks)."""

    def __init__(self, model, total_steps, lr, weight_decay, warmup_ratio, max_grad_norm=1.0, monitor_ntp_acc=True):
        self.model, self.monitor_ntp_acc, self.max_grad_norm = model, monitor_ntp_acc, max_grad_norm
        self.opt = torch.optim.AdamW(hf_param_groups(model, weight_decay), lr=lr, betas=(0.9, 0.999), eps=1e-8)
        warm = math.ceil(total_steps * warmup_ratio)

        def lam(step):
            if step < warm:
                return step / max(1, warm)
            return max(0.0, 0.5 * (1.0 + math.cos(math.pi * (step - warm) / max(1, total_steps - warm))))
        self.sched = torch.optim.lr_scheduler.LambdaLR(self.opt, lam)
        self.logs = []

    def compute_loss(self, model, inputs, return_outputs=False, greedy=None):
        outputs = model(**inputs)
        if model.training and self.monitor_ntp_acc and 'labels' in inputs:
            # train_util_wrap.py:106: `outputs.logits.detach().argmax(axis=-1)`; in training mode the reference's forward returns no logits
            # (transformer_xl.py:194), so the greedy ids come from `greedy(outputs)`: the LM-head kernel's fused arg-max / an eval forward
            preds = greedy(outputs)
            labels_ = inputs['labels'].detach()
            preds, labels_ = preds[:, :-1], labels_[:, 1:]
            msk = labels_ != PT_LOSS_PAD
            self.logs.append(dict(ntp_acc=(preds[msk] == labels_[msk]).sum().item() / int(msk.sum().item())))
        if isinstance(outputs, dict) and 'loss' not in outputs:
            raise ValueError('The model did not return a loss from the inputs')
        loss = outputs['loss'] if isinstance(outputs, dict) else outputs[0]
        return (loss, outputs) if return_outputs else loss

    def training_step(self, inputs, greedy):
        self.model.train()
        loss = self.compute_loss(self.model, inputs, greedy=greedy)
        loss.backward()
        gn = torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.max_grad_norm)
        lr = self.opt.param_groups[0]['lr']
        self.opt.step()
        self.sched.step()
        self.model.zero_grad()
        self.logs[-1].update(loss=float(loss.detach()), grad_norm=float(gn), learning_rate=lr)
        return loss.detach()

    def prediction_step(self, inputs):
        """trainer_eval_wrap.py: loss, then every output that is not ignored / not the loss, then max_out_logits."""
        self.model.eval()
        with torch.no_grad():
            loss, outputs = self.compute_loss(self.model, inputs, return_outputs=True)
        assert isinstance(outputs, dict)
        logits = tuple(v for k, v in outputs.items() if k not in IGNORE_KEYS_FOR_EVAL + ['loss'])
        assert len(logits) == 1                              # only prediction_scores survives the ignore list
        preds = logits[0].argmax(dim=-1)                     # preprocess_logits_for_metrics = max_out_logits
        return loss.mean().detach(), preds, inputs['labels'], inputs['key_scores'].detach()


def _ntp_acc(preds, labels):                                 # ComputeMetrics.__call__ with clm_pred_shifted=False
    preds, labels = preds[:, :-1].flatten(), labels[:, 1:].flatten()
    msk = labels != PT_LOSS_PAD
    return (preds[msk] == labels[msk]).float().mean().item()


@pytest.mark.gpu
@pytest.mark.parametrize('cutoffs', [[], [300]])
def test_model_under_the_reference_trainer_hooks(pkg, tmp_path, cutoffs):
    ref, model = make_pair(pkg, 'fp32', cutoffs=cutoffs)
    g = torch.Generator().manual_seed(77)
    batches = []
    for _ in range(4):
        ids = torch.randint(2, 422, (3, 32), generator=g)
        ids[1, 24:] = 1
        labels = torch.where(ids == 1, torch.full_like(ids, PT_LOSS_PAD), ids)
        batches.append(dict(input_ids=ids, labels=labels, key_scores=torch.rand(3, 24, generator=g)))
    kw = dict(total_steps=3, lr=1e-3, weight_decay=1e-2, warmup_ratio=0.34)
    tr_ref, tr = _MyTrainerLike(ref, **kw), _MyTrainerLike(model, **kw)
    model.monitor_greedy = True

    def ref_greedy(_outputs, inputs=None):
        ref.eval()
        with torch.no_grad():
            p = ref(input_ids=cur['input_ids']).logits.argmax(-1)
        ref.train()
        return p

    for step in range(3):
        cur = {k: v.clone() for k, v in batches[step].items()}
        tr_ref.training_step(cur, ref_greedy)
        dev = {k: v.cuda() for k, v in batches[step].items()}
        tr.training_step(dev, lambda _o: model.last_greedy)
        a, b = tr.logs[-1], tr_ref.logs[-1]
        assert abs(a['learning_rate'] - b['learning_rate']) < 1e-12
        assert abs(a['loss'] - b['loss']) / b['loss'] < 2e-4, (a, b)
        assert abs(a['grad_norm'] - b['grad_norm']) / b['grad_norm'] < 2e-3, (a, b)
        assert abs(a['ntp_acc'] - b['ntp_acc']) < 0.03, (a, b)
        if step == 1:      # checkpoint round trip in the middle of training (train.py:591 / eval.py:91); the optimiser keeps the old module
            model.save_pretrained(tmp_path)
            again = pkg.MyTransfoXLLMHeadModel.from_pretrained(tmp_path, compute_dtype='fp32').cuda()
            for k, v in model.state_dict().items():
                assert torch.equal(again.state_dict()[k].cpu(), v.cpu()), k
    # parameters after three stock-AdamW steps == the oracle's
    got = dict(model.named_parameters())
    for name, p in ref.named_parameters():
        assert torch.allclose(got[name].detach().cpu(), p.detach(), rtol=2e-4, atol=3e-5), name      # Adam amplifies fp32 noise where |g| ~ eps
    # evaluation through prediction_step: loss, greedy ids, labels, key_scores
    ev = batches[3]
    l_r, p_r, lab_r, ks_r = tr_ref.prediction_step({k: v.clone() for k, v in ev.items()})
    l_g, p_g, lab_g, ks_g = tr.prediction_step({k: v.cuda() for k, v in ev.items()})
    assert abs(l_g.item() - l_r.item()) / l_r.item() < 2e-4
    assert (p_g.cpu() != p_r).float().mean().item() < 0.02          # near-ties of two log-probs aside
    assert abs(_ntp_acc(p_g.cpu(), lab_g.cpu()) - _ntp_acc(p_r, lab_r)) < 0.03
    assert torch.equal(ks_g.cpu(), ks_r)
