"""Adam / ExpDecay / EMA mirrors (optims/{adam,expdecay,ema}.py) on top of ONE fused kernel (ngp_adam_ema).

Reference step order (runner/runner.py:75-76): optimizer.step(loss) [ExpDecay -> jt.nn.Adam] then
ema_optimizer.ema_step(), which overwrites the live parameters with the debiased EMA (ema.py:26-37).
Here `Adam.step` runs backward and, when an EMA is attached to the same parameters, defers the parameter update so
that `EMA.ema_step` can apply Adam + EMA + gradient zeroing in a single streaming pass.  Without an EMA the same
kernel runs with decay 0 (pure Adam)."""
import torch

from .. import ops
from ..utils.registry import OPTIMS


class _ParamState:
    def __init__(self, p):
        self.p = p
        self.m = torch.zeros(p.numel(), dtype=torch.float32, device=p.device)
        self.v = torch.zeros(p.numel(), dtype=torch.float32, device=p.device)
        self.master = p.detach().float().reshape(-1).clone()     # EMA `values` (ema.py:16-19) == fp32 master copy


@OPTIMS.register_module()
class Adam:
    def __init__(self, params, lr=1e-1, eps=1e-15, betas=(0.9, 0.99), weight_decay=0):
        assert weight_decay == 0, "the NGP configs use weight_decay = 0"
        self.params = [p for p in params if p.requires_grad]
        self.lr, self.eps, self.betas = lr, eps, betas
        self.n_step = 0
        self.state = [_ParamState(p) for p in self.params]
        self._ema = None
        self._pending = False

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    def backward(self, loss):
        loss.sum().backward()                    # Jittor differentiates an unreduced loss as its sum (runner.py:74-75)

    def step(self, loss=None):
        if loss is not None:
            self.backward(loss)
        self.n_step += 1
        if self._ema is not None:
            self._pending = True                 # EMA.ema_step() applies the fused update
        else:
            self._apply(ema_decay=0.0)

    def _apply(self, ema_decay, grad_scale=1.0):
        for st in self.state:
            g = st.p.grad
            if g is None:
                continue
            g = g.reshape(-1)
            if st.p.dtype == torch.float32 and g.dtype != torch.float32:
                g = g.float()
            ops.adam_ema(st.p.data.view(-1), g, st.m, st.v, st.master, self.lr, self.n_step, self.betas[0], self.betas[1], self.eps,
                         ema_decay, grad_scale=grad_scale, zero_grad=False)
            st.p.grad = None
        self._pending = False

    def state_dict(self):
        return {"n_step": self.n_step, "lr": self.lr, "m": [s.m for s in self.state], "v": [s.v for s in self.state],
                "master": [s.master for s in self.state]}

    def load_state_dict(self, sd):
        self.n_step, self.lr = sd["n_step"], sd["lr"]
        for s, m, v, ms in zip(self.state, sd["m"], sd["v"], sd["master"]):
            s.m.copy_(m); s.v.copy_(v); s.master.copy_(ms)


@OPTIMS.register_module()
class ExpDecay:
    """optims/expdecay.py:7-31: lr *= decay_base every decay_interval steps from decay_start on."""

    def __init__(self, nested_optimizer, decay_start, decay_interval, decay_base, decay_end=None):
        self.base_lr = nested_optimizer.lr
        self._nested_optimizer = nested_optimizer
        self.decay_start, self.decay_interval, self.decay_base = decay_start, decay_interval, decay_base
        self.decay_end = 10000000 if decay_end is None else decay_end
        self.steps = 0
        self.m_learning_rate_factor = 1

    def advance_lr(self):
        if self.steps >= self.decay_start and (self.steps - self.decay_start) % self.decay_interval == 0 and self.steps <= self.decay_end:
            self.m_learning_rate_factor *= self.decay_base
        self._nested_optimizer.lr = self.base_lr * self.m_learning_rate_factor
        self.steps += 1
        return self._nested_optimizer.lr

    def step(self, loss=None):
        self.advance_lr()
        self._nested_optimizer.step(loss)

    def zero_grad(self):
        return self._nested_optimizer.zero_grad()

    def state_dict(self):
        return {"steps": self.steps, "m_learning_rate_factor": self.m_learning_rate_factor}

    def load_state_dict(self, sd):
        self.steps, self.m_learning_rate_factor = sd["steps"], sd["m_learning_rate_factor"]


@OPTIMS.register_module()
class EMA:
    """optims/ema.py:7-38."""

    def __init__(self, params, decay, adam=None):
        self.decay = decay
        self.steps = 0
        self.params = [p for p in params if p.requires_grad]
        self._adam = adam
        if adam is not None:
            adam._ema = self

    def attach(self, adam):
        inner = getattr(adam, "_nested_optimizer", adam)
        assert [id(p) for p in inner.params] == [id(p) for p in self.params], "EMA and Adam must cover the same parameters"
        self._adam = inner
        inner._ema = self

    def ema_step(self, loss=None):
        assert loss is None
        self.steps += 1
        assert self._adam is not None and self._adam._pending, "EMA.ema_step must follow optimizer.step (runner.py:75-76)"
        assert self.steps == self._adam.n_step
        self._adam._apply(ema_decay=self.decay)

    def state_dict(self):
        return {"steps": self.steps, "decay": self.decay}

    def load_state_dict(self, sd):
        self.steps = sd["steps"]
