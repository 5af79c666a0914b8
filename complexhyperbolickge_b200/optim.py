"""Caller-side contract of the hot path: negative sampling + loss + epoch loop.

The reference's ``optimizers/kg_optimizer.py:KGOptimizer`` is kept UNCHANGED as the contract (SURVEY §8a
row L): ``get_neg_samples(batch) -> LongTensor (B, neg)`` of tail ids != true tail (:92-99),
``neg_sampling_loss`` = -mean(cat[logsigmoid(s_pos), logsigmoid(-s_neg)]) over B(1+neg) terms with two
``model(queries, tails)`` calls (:101-123), ``calculate_loss`` adds the regulariser on the positive call's
factors (:174-197), ``epoch`` shuffles, steps every ``update_steps`` batches and returns the mean of the
per-batch losses (:239-277).  The reference class itself can drive our models directly; this file restates
the same contract so the drop-in loop also runs where /root/reference is absent (the GPU box), and adds the
data-parallel variant north_star asks for (sparse row-gradient exchange is SURVEY §8f "next").
Only the negative-sampling branch is in scope (neg_sample_size > 0).
"""
import torch
import torch.nn.functional as F


class N3(torch.nn.Module):
    """optimizers/regularizers.py:45-58."""

    def __init__(self, weight: float):
        super().__init__()
        self.weight = weight

    def forward(self, factors):
        norm = 0
        for f in factors:
            norm += self.weight * torch.sum(torch.abs(f) ** 3)
        return norm / factors[0].shape[0]


class F2(torch.nn.Module):
    """optimizers/regularizers.py:21-30."""

    def __init__(self, weight: float):
        super().__init__()
        self.weight = weight

    def forward(self, factors):
        norm = 0
        for f in factors:
            norm += self.weight * torch.sum(f ** 2)
        return norm / factors[0].shape[0]


class KGOptimizer(object):
    def __init__(self, model, regularizer, optimizer, batch_size, update_steps, neg_sample_size, double_neg,
                 optimizer2=None, loss="crossentropy", smoothing=None, verbose=True):
        if neg_sample_size <= 0:
            raise NotImplementedError("only the negative-sampling loss is on the accelerated path (SURVEY §2 #5)")
        self.model, self.regularizer, self.optimizer = model, regularizer, optimizer
        self.optimizer.zero_grad()
        self.batch_size, self.update_steps, self.verbose = batch_size, update_steps, verbose
        # The reference stores double_neg and ignores it at HEAD; the sampler that used it is still there,
        # commented out (:78-91): every negative replaces the tail AND, with double_neg, the head.  Restored with those
        # semantics: negative j of triple b is (h'_bj, r_b, t'_bj), scored with per-pair queries of shape (B, neg, 2).
        self.double_neg = bool(double_neg)
        self.neg_sample_size = neg_sample_size
        self.n_entities = model.sizes[0]
        self.device = model.entity.weight.device

    def get_neg_samples(self, input_batch):
        negsamples = torch.randint(0, self.n_entities - 1, size=(input_batch.shape[0], self.neg_sample_size),
                                   device=input_batch.device)
        return torch.where(negsamples < input_batch[:, 2].unsqueeze(-1), negsamples, negsamples + 1)

    def get_neg_heads(self, input_batch):
        """double_neg: corrupted head per negative, uniform over the entities != true head (mirror of get_neg_samples)."""
        neg = torch.randint(0, self.n_entities - 1, size=(input_batch.shape[0], self.neg_sample_size), device=input_batch.device)
        return torch.where(neg < input_batch[:, 0].unsqueeze(-1), neg, neg + 1)

    def neg_sampling_loss(self, input_batch):
        positive_score, factors = self.model(input_batch[:, :2].unsqueeze(1), input_batch[:, 2].unsqueeze(1))
        positive_score = F.logsigmoid(positive_score)
        neg_samples = self.get_neg_samples(input_batch)
        neg_queries = input_batch[:, :2].unsqueeze(1)
        if self.double_neg:
            neg_heads = self.get_neg_heads(input_batch)
            neg_queries = torch.stack([neg_heads, input_batch[:, 1:2].expand_as(neg_heads)], -1)
        negative_score, _ = self.model(neg_queries, neg_samples)
        negative_score = F.logsigmoid(-negative_score)
        loss = -torch.cat([positive_score.view(-1), negative_score.view(-1)]).mean()
        return loss, factors

    def calculate_loss(self, input_batch):
        loss, factors = self.neg_sampling_loss(input_batch)
        loss += self.regularizer.forward(factors)
        return loss

    def calculate_valid_loss(self, examples):
        b_begin, loss, counter = 0, 0.0, 0
        with torch.no_grad():
            while b_begin < examples.shape[0]:
                input_batch = examples[b_begin:b_begin + self.batch_size].to(self.device)
                b_begin += self.batch_size
                loss += self.calculate_loss(input_batch)
                counter += 1
        return loss / counter

    def epoch(self, examples):
        actual_examples = examples[torch.randperm(examples.shape[0]), :]
        b_begin, total_loss, counter = 0, 0.0, 0
        while b_begin < examples.shape[0]:
            input_batch = actual_examples[b_begin:b_begin + self.batch_size].to(self.device)
            l = self.calculate_loss(input_batch)
            l.backward()
            if self.update_steps == 1 or (counter + 1) % self.update_steps == 0 or \
                    b_begin + self.batch_size >= examples.shape[0]:
                self.optimizer.step()
                self.optimizer.zero_grad()
            b_begin += self.batch_size
            total_loss += l.item()
            counter += 1
        return total_loss / counter
