"""Process-wide "weights may have changed" counter.

Plans keep packed bf16 copies of the fp32 nn.Parameters and re-pack them when the parameters change.  The
tensors' autograd version counters catch load_state_dict / foreach optimizers / in-place edits, but NOT
every writer: torch's fused optimizers (Adam(fused=True)) update parameters without bumping `_version`,
and the train plans write BatchNorm running statistics through raw pointers.  So every hand-written
backward pass bumps this counter -- an optimizer step may follow it -- and every plan folds it into its
weight-version key: a training loop re-packs once per step whatever the optimizer, an eval plan re-packs
once after training and then stays on its cached operands."""
_epoch = 0


def note_backward() -> None:
    global _epoch
    _epoch += 1


def value() -> int:
    return _epoch


# Training plans with batched weight packing register here when they are built, so that the fused optimizers
# (rtsds_b200/optim.py) can refresh their packed operands inside the optimizer kernel.  Weak references: a plan
# dropped with its model disappears from the list.
import weakref

_live = []


def register_plan(plan) -> None:
    _live.append(weakref.ref(plan))


def live_plans():
    global _live
    out = [r() for r in _live]
    _live = [r for r, p in zip(_live, out) if p is not None]
    return [p for p in out if p is not None]


class PlanOwner:
    """Mixin of the drop-in top-level modules (BiSeNet, ResNetMulti, the discriminators).  Execution plans — and the CUDA
    graphs captured from them — hold RAW POINTERS into the module's parameters and buffers (ARM / FFM / bias vectors and
    BatchNorm running statistics are read in place).  Anything that gives the module NEW storage invalidates them:
    `.to()/.cuda()/.cpu()/.float()` (all go through `_apply`) and `load_state_dict` (assign=True swaps the Parameter
    objects; the plain form keeps storage but is rare enough to simply rebuild too).  Both drop every cached plan, so the
    next forward rebuilds its launch list and re-captures its graph from the live tensors.  Re-pointing `param.data` by
    hand is not detected — call `model.rtsds_drop_plans()` after doing that."""

    def rtsds_drop_plans(self):
        for k in ("_rtsds_plans", "_rtsds_train_plans", "_rtsds_bn_counters"):
            self.__dict__.pop(k, None)

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.rtsds_drop_plans()
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.rtsds_drop_plans()
        return out
