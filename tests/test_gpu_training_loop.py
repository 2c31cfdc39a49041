"""The loss section of the reference's batch loop (main_deep_staple.py:701-775) run twice on the same
tiny 3D network, data and seeds: once with the reference's own lines executed on CUDA tensors (stock ATen
kernels), once with deep_staple_b200 substituted at the call sites exactly as INTEGRATION.md describes.
Both the out-of-line mode (:723-729) and the in-line mode (:731-736) are covered, under CUDA autocast
with two GradScalers, AdamW for the network and SparseAdam for the data parameters."""
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu


class TinySeg3D(nn.Module):  # stands in for MobileNet_LR_ASPP_3D: returns {'out': logits [B, C, D, H, W]}
    def __init__(self, classes=2):
        super().__init__()
        self.a = nn.Conv3d(1, 8, 3, padding=1)
        self.b = nn.Conv3d(8, classes, 1)

    def forward(self, x):
        return {"out": self.b(torch.relu(self.a(x)))}


def _run(use_ds: bool, ool: bool, dev, steps=4):
    import deep_staple_b200 as ds

    torch.manual_seed(0)
    M, B, S = 24, 4, (16, 16, 16)
    lraspp = TinySeg3D().to(dev)
    optimizer = torch.optim.AdamW(lraspp.parameters(), lr=1e-3)
    embedding = nn.Embedding(M, 1, sparse=True).to(dev)
    torch.nn.init.normal_(embedding.weight.data, mean=0.0, std=0.0)  # :438
    optimizer_dp = torch.optim.SparseAdam(embedding.parameters(), lr=0.1, betas=(0.9, 0.999), eps=1e-8)  # :442
    scaler, scaler_dp = torch.amp.GradScaler("cuda"), torch.amp.GradScaler("cuda")  # :407, :445
    gen = torch.Generator().manual_seed(1)
    images = torch.randn((M, 1, *S), generator=gen)
    labels = (torch.rand((M, *S), generator=gen) < 0.3).long()
    gt_num = labels.reshape(M, -1).sum(1).float()
    fixed_weighting = ((gt_num + torch.e).log() + torch.e).to(dev)  # :653
    counts = torch.bincount(labels.reshape(-1), minlength=2).float()
    class_weights = 1 / counts.pow(0.35)
    class_weights = (class_weights / class_weights.mean()).to(dev)  # :650-651
    n_dims = (-3, -2, -1)
    losses = []
    for step in range(steps):
        b_idxs_dataset = torch.arange(step * B, (step + 1) * B) % M
        b_img = images[b_idxs_dataset].to(dev)
        b_seg_modified = labels[b_idxs_dataset].to(dev)
        b_idxs_dataset = b_idxs_dataset.to(dev)
        optimizer.zero_grad()
        optimizer_dp.zero_grad()
        with torch.autocast("cuda", enabled=True):  # :701
            logits = lraspp(b_img)["out"]  # :708 (fp16)
            if use_ds:
                ce_loss = ds.weighted_ce(logits, b_seg_modified, class_weights, grad_scaler=scaler)
            else:
                ce_loss = nn.CrossEntropyLoss(class_weights)(logits, b_seg_modified)  # :716
            if ool:
                scaler.scale(ce_loss).backward()  # :719
                scaler.step(optimizer)
                scaler.update()
                for p in lraspp.parameters():
                    p.requires_grad = False
                dp_logits = lraspp(b_img)["out"]  # :727 second forward with the stepped weights
                for p in lraspp.parameters():
                    p.requires_grad = True
            else:
                dp_logits = logits  # :733
            if use_ds:
                dp_loss = ds.dp_weighted_loss(dp_logits, b_seg_modified, embedding(b_idxs_dataset).squeeze(),
                                              fixed_weighting[b_idxs_dataset], True, grad_scaler=scaler_dp)
            else:
                dp_loss = nn.CrossEntropyLoss(reduction="none")(dp_logits, b_seg_modified)  # :738
                dp_loss = dp_loss.mean(n_dims)
                bare_weight = embedding(b_idxs_dataset).squeeze()
                weight = torch.sigmoid(bare_weight)
                weight = weight / weight.mean()
                weight = weight / fixed_weighting[b_idxs_dataset]
                p_pred_num = (dp_logits.argmax(1) > 0).sum(dim=n_dims).detach()
                risk_regularization = -weight * p_pred_num / (dp_logits.shape[-3] * dp_logits.shape[-2] * dp_logits.shape[-1])
                dp_loss = (dp_loss * weight).sum() + risk_regularization.sum()  # :757
        scaler_dp.scale(dp_loss).backward()  # :762
        if not ool:
            scaler_dp.step(optimizer)  # :765
        scaler_dp.step(optimizer_dp)  # :767
        scaler_dp.update()
        losses.append(float(dp_loss.item()))  # :775
    return losses, embedding.weight.detach().clone(), [p.detach().clone() for p in lraspp.parameters()]


@pytest.mark.parametrize("ool", [True, False])
def test_reference_loop_lines_vs_substituted_call_sites(cuda_device, ool):
    ref_losses, ref_emb, ref_params = _run(False, ool, cuda_device)
    our_losses, our_emb, our_params = _run(True, ool, cuda_device)
    for a, b in zip(our_losses, ref_losses):
        assert abs(a - b) <= 2e-3 * abs(b) + 1e-4, (our_losses, ref_losses)  # fp16 network, 4 optimiser steps
    touched = (ref_emb != 0).reshape(-1)
    assert torch.equal(touched, (our_emb != 0).reshape(-1))  # the same rows were updated
    assert torch.allclose(our_emb, ref_emb, rtol=2e-2, atol=2e-3)
    for a, b in zip(our_params, ref_params):
        assert torch.allclose(a, b, rtol=1e-2, atol=1e-4)
