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
