"""The two helpers of the reference's util.py that the MCD hot path consumes (util.py:99-111 class weights,
util.py:44-48 prediction entropy) plus the LR schedule (util.py:87-96).  The rest of that file (check-pointing,
json dumps, colourised label PNGs, interactive prompts) is host I/O outside SURVEY.md section 8."""
import torch

from mcd_b200 import ops


def get_class_weight_from_file(n_class, weight_filename=None, add_bg_loss=False):
    """ones(n_class) (optionally scaled by a csv with columns class_id, weight); the background class
    n_class-1 gets weight 0 unless add_bg_loss."""
    weight = torch.ones(n_class)
    if weight_filename:
        import pandas as pd
        loss_df = pd.read_csv(weight_filename)
        loss_df.sort_values("class_id", inplace=True)
        weight *= torch.FloatTensor(loss_df.weight.values)
    if not add_bg_loss:
        weight[n_class - 1] = 0
    return weight


def calc_entropy(output):
    """-mean(p * log(p + 1e-6)), p = softmax(output, dim=1): one fused kernel over the logits."""
    logits = output if output.dtype in (torch.bfloat16, torch.float32) else output.float()
    _, ent = ops.argmax_entropy(logits.contiguous(), want_labels=False, want_entropy=True)
    return ent


def predict_labels(output, n_valid_class=None):
    """argmax over the first n_valid_class channels (testers drop the background channel:
    adapt_tester.py:121-124, adapt_triple_multitask_tester.py:139-142) -> int64 [B,H,W]."""
    logits = output if output.dtype in (torch.bfloat16, torch.float32) else output.float()
    labels, _ = ops.argmax_entropy(logits.contiguous(), c_arg=n_valid_class, want_labels=True,
                                   want_entropy=False)
    return labels


def adjust_learning_rate(optimizer, lr_init, decay_rate, epoch, num_epochs):
    """step decay at 1/2 and 3/4 of the schedule (the trainers pass weight_decay as decay_rate,
    adapt_trainer.py:228-230); written into every param group."""
    lr = lr_init
    if epoch >= num_epochs * 0.75:
        lr *= decay_rate ** 2
    elif epoch >= num_epochs * 0.5:
        lr *= decay_rate
    for param_group in optimizer.param_groups:
        param_group['lr'] = lr
    return lr
