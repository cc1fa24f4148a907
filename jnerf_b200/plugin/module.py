"""Jittor-style module base: `execute` is the forward method (jt.nn.Module convention), parameters are torch Parameters."""
import torch


class Module(torch.nn.Module):
    def forward(self, *args, **kwargs):
        return self.execute(*args, **kwargs)

    def execute(self, *args, **kwargs):
        raise NotImplementedError
