"""GPU: labels uploaded through ds.LabelUploader (packed to uint8 on the host) give the same results, bit for bit,
as the reference's int64 upload (main_deep_staple.py:687) -- loss path and consensus vote."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_label_uploader_loss_identical_to_int64_upload(cuda_device):
    import deep_staple_b200 as ds

    gen = torch.Generator().manual_seed(0)
    B, C, S = 4, 3, (16, 24, 32)
    up = ds.LabelUploader((B, *S), cuda_device)
    for step in range(5):  # more steps than staging slots: slot reuse is exercised
        logits = torch.randn((B, C, *S), generator=gen)
        target = torch.randint(0, C, (B, *S), generator=gen)
        bare = torch.randn(B, generator=gen)
        fixed = torch.rand(B, generator=gen) * 5 + 1
        outs = []
        for lab in (target.to(cuda_device), up(target.pin_memory())):
            z = logits.to(cuda_device).requires_grad_(True)
            w = bare.to(cuda_device).requires_grad_(True)
            loss = ds.dp_weighted_loss(z, lab, w, fixed.to(cuda_device), True)
            loss.backward()
            outs.append((loss.detach().clone(), z.grad.clone(), w.grad.clone()))
        assert up(target).dtype == torch.uint8
        for a, b in zip(*outs):
            assert torch.equal(a, b), f"step {step}: packed-label upload changed the result"


def test_label_uploader_vote_identical_and_refuses_out_of_range(cuda_device):
    import deep_staple_b200 as ds

    gen = torch.Generator().manual_seed(2)
    N, S = 10, (16, 16, 32)
    labels = (torch.rand((N, *S), generator=gen) < 0.4).long()
    dps = [float(x) for x in torch.randn(N, generator=gen)]
    up = ds.LabelUploader((N, *S), cuda_device)
    ref = ds.calc_dp_consensus(list(labels.to(cuda_device)), dps)
    out = ds.calc_dp_consensus(list(up(labels)), dps)
    assert out.dtype == torch.int64 and torch.equal(out, ref)  # calc_dp_consensus returns int64 like the reference
    labels[3, 1, 2, 3] = 1000
    with pytest.raises(ValueError, match="outside"):
        up(labels)


def test_calc_dp_consensus_from_host_volumes_takes_the_packed_path(cuda_device):
    """the notebook's call with DEVICE='cpu' (consensus.ipynb:30): int64 host volumes in, int64 host map out"""
    import deep_staple_b200 as ds
    from oracle import consensus as oc

    gen = torch.Generator().manual_seed(4)
    N, S = 10, (12, 16, 20)
    labels = (torch.rand((N, *S), generator=gen) < 0.4).long()
    dps = [float(x) for x in torch.randn(N, generator=gen)]
    for _ in range(3):  # the pinned staging slot is reused
        out = ds.calc_dp_consensus(list(labels), dps)
        assert out.device.type == "cpu" and out.dtype == torch.int64
        assert torch.equal(out, ds.calc_dp_consensus(list(labels.to(cuda_device)), dps).cpu())
    ref = oc.calc_dp_consensus(list(labels), dps)
    near = oc.near_tie_mask(list(labels), dps)
    assert not bool(((out != ref) & ~near).any())
    # label values above 255 cannot be packed: the call falls back to the int64 upload and still multiplies VALUES
    big = labels.clone()
    big[0] *= 300
    out_big = ds.calc_dp_consensus(list(big), dps)
    assert torch.equal(out_big, ds.calc_dp_consensus(list(big.to(cuda_device)), dps).cpu())


def test_label_uploader_takes_labels_packed_earlier(cuda_device):
    """uint8 host labels (packed by ``narrow_labels_host`` in a data-loader worker) are copied as they are: pinned,
    pageable and non-contiguous sources, over more steps than there are staging slots; wrong shapes are refused."""
    import deep_staple_b200 as ds

    gen = torch.Generator().manual_seed(4)
    shape = (3, 10, 12, 16)
    up = ds.LabelUploader(shape, cuda_device)
    for step in range(5):
        target = torch.randint(0, 4, shape, generator=gen)
        packed = ds.narrow_labels_host(target)
        ref = target.to(torch.uint8)
        for src in (packed, packed.pin_memory(), packed.transpose(1, 2).contiguous().transpose(1, 2)):
            got = up(src)
            assert got.dtype == torch.uint8 and got.device.type == "cuda"
            assert torch.equal(got.cpu(), ref), f"step {step}"
        assert torch.equal(up(target).cpu(), ref)  # the int64 route through the same slots
    with pytest.raises(ValueError):
        up(torch.zeros((3, 10, 12, 15), dtype=torch.uint8))
