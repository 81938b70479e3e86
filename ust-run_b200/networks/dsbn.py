"""Domain-specific BatchNorm (drop-in for the reference's networks/dsbn.py:4-34).

A ModuleList ``bns`` of ``num_domains`` nn.BatchNorm2d parameter holders; ``forward(x, domain_label)``
sends the WHOLE batch through ``bns[domain_label[0]]`` (0-based) and returns ``(y, domain_label)``.
Inside a network the selected BN is consumed by the fused conv/BN kernels (``select``); called on
its own the module runs the stand-alone BN program below (per-channel statistics by the two-stage
reduction kernel, finalize, apply) -- still on the sm_100a library, never on torch's BatchNorm.
"""
import torch
from torch import nn

from ustrun import engine as E
from ustrun.bridge import run_program


class _DomainSpecificBatchNorm(nn.Module):
    _version = 2

    def __init__(self, num_features, num_domains, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super(_DomainSpecificBatchNorm, self).__init__()
        self.bns = nn.ModuleList(
            nn.BatchNorm2d(num_features, eps, momentum, affine, track_running_stats) for _ in range(num_domains))

    def reset_running_stats(self):
        for bn in self.bns:
            bn.reset_running_stats()

    def reset_parameters(self):
        for bn in self.bns:
            bn.reset_parameters()

    def _check_input_dim(self, input):
        raise NotImplementedError

    def select(self, domain_label):
        """The nn.BatchNorm2d this batch uses (indexing with a GPU tensor syncs, as upstream does)."""
        return self.bns[int(domain_label[0])]

    def forward(self, x, domain_label):
        self._check_input_dim(x)
        bn = self.select(domain_label)
        y = run_program(bn, lambda ctx, a: (E.batchnorm_only(ctx, a, bn),), x, training=self.training)[0]
        return y, domain_label


class DomainSpecificBatchNorm2d(_DomainSpecificBatchNorm):
    def _check_input_dim(self, input):
        if input.dim() != 4:
            raise ValueError('expected 4D input (got {}D input)'.format(input.dim()))
