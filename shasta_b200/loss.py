"""The reference's training loss, restated (tools/nusc_shasta/train.py:201-211): two masked cross-entropy terms on
the forward (matched1) and backward (matched2) affinities against the augmented ground-truth matrix gt (B,M+2,M+2),
each normalised by its number of positives (left un-normalised when there is none). Plain torch ops on the kernel
outputs, exactly like the reference's training loop."""
import torch


def affinity_loss(matched1, matched2, gt, eps=1e-10):
    gt1 = gt[:, :-2, :]
    gt2 = gt[:, :, :-2]
    loss_f = torch.mul(gt1, -torch.log(matched1 + eps)).sum()
    if gt1.sum() > 0:
        loss_f = loss_f / gt1.sum()
    loss_b = torch.mul(gt2, -torch.log(matched2 + eps)).sum()
    if gt2.sum() > 0:
        loss_b = loss_b / gt2.sum()
    return (loss_f + loss_b) / 2.0
