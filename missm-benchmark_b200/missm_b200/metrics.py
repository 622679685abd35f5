"""Device-resident evaluation metrics (csrc/metrics.cu).

`EvalAccumulator` replaces the per-batch host traffic of the reference's evaluate() (train_ddp.py:88-133,
test.py:21-66): `update(outputs, labels)` is one asynchronous launch per batch, `compute()` is the epoch's only
synchronisation and returns the reference's dictionary {'loss', 'accuracy', 'f1', 'auc'} -- accuracy and macro-F1
from the confusion matrix (the definitions of sklearn.metrics.accuracy_score / f1_score(average='macro')), AUC from
the kept softmax scores through sklearn's roc_auc_score(multi_class='ovo') exactly as the script calls it.  Under
data parallelism the script all-gathers predictions from every rank into every rank; here every rank accumulates
its own shard and `compute(all_reduce=True)` sums the accumulators (one small all-reduce)."""
import ctypes

import torch

from ._lib import check, lib, stream_ptr


class EvalAccumulator:
    def __init__(self, n_classes, device):
        self.C, self.device = int(n_classes), torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError("missm_b200: EvalAccumulator lives on a CUDA device (no CPU fallback)")
        self.confusion = torch.zeros((self.C, self.C), device=self.device, dtype=torch.int64)
        self.loss_sum = torch.zeros((1,), device=self.device, dtype=torch.float64)
        self.n_seen = torch.zeros((1,), device=self.device, dtype=torch.int64)
        self.n_batches = 0
        self._probs, self._labels = [], []

    def update(self, outputs, labels):
        """outputs: logits [B, C] (CUDA); labels: int64 [B].  Asynchronous."""
        if not outputs.is_cuda:
            raise RuntimeError("missm_b200: EvalAccumulator.update needs CUDA logits (no CPU fallback)")
        logits = outputs.detach().float().contiguous()
        labels = labels.detach().to(device=logits.device, dtype=torch.int64).contiguous()
        B, C = logits.shape
        assert C == self.C, (C, self.C)
        probs = torch.empty_like(logits)
        p = lambda t: ctypes.c_void_p(t.data_ptr())      # noqa: E731
        check(lib().missm_eval_accumulate(p(logits), p(labels), B, C, p(probs), p(self.confusion), p(self.loss_sum),
                                          p(self.n_seen), stream_ptr()), "eval_accumulate")
        self.n_batches += 1
        self._probs.append(probs)
        self._labels.append(labels)

    def compute(self, all_reduce=False, with_auc=True):
        conf, loss, nb = self.confusion.clone(), self.loss_sum.clone(), float(self.n_batches)
        if all_reduce:
            import torch.distributed as dist
            nbt = torch.tensor([nb], device=self.device, dtype=torch.float64)
            for t in (conf, loss, nbt):
                dist.all_reduce(t)
            nb = float(nbt)
        conf = conf.cpu().double()                        # the epoch's read-out: rows = true label, columns = prediction
        total = conf.sum().clamp_min(1.0)
        tp = conf.diag()
        pred_n, true_n = conf.sum(0), conf.sum(1)
        present = (pred_n + true_n) > 0                   # sklearn averages over the labels that occur in y_true or y_pred
        f1 = torch.where(pred_n + true_n > 0, 2 * tp / (pred_n + true_n).clamp_min(1.0), torch.zeros_like(tp))
        out = {'loss': float(loss.cpu()) / max(nb, 1.0), 'accuracy': float(tp.sum() / total),
               'f1': float(f1[present].mean()) if present.any() else 0.0}
        if with_auc:
            from sklearn.metrics import roc_auc_score
            probs = torch.cat(self._probs).cpu().numpy()
            labels = torch.cat(self._labels).cpu().numpy()
            out['auc'] = float(roc_auc_score(labels, probs if self.C > 2 else probs[:, 1], multi_class='ovo'))
        return out
