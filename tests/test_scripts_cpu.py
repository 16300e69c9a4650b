"""Host-side checks of the script-level harness that need no GPU: the fabricated dataset is what the reference's own
`ReferenceDataset` / `get_reference_dataloader` (dataloader.py:19-46, 122-266) read, the metric shim offers the classes the
scripts import, and the offline stubs build the loss networks without a download."""
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from baseline import reference as R  # noqa: E402

needs_ref = pytest.mark.skipif(not R.available(), reason="baseline/_ref (copy of the reference) is not present")


@needs_ref
def test_fabricated_dataset_loads_through_the_reference_dataloader(tmp_path):
    info = R.fabricate_dataset(tmp_path, n_ids=3, per_id=2, full=256)
    from face_mask_inpaint_b200.patch import _ensure_msssim_shim
    _ensure_msssim_shim()
    sys.path.insert(0, str(R.REF))
    try:
        import dataloader
        ds = dataloader.ReferenceDataset(tmp_path / "src", tmp_path / "ref", tmp_path / "mask", tmp_path / "identity.txt",
                                         apply_transform=False, scale=0.25, use_ssim=False, return_id=True)
        assert len(ds) == info["n"] == 6
        item = ds[0]
        assert item["src_img"].shape == (3, 64, 64) and item["mask"].shape == (64, 64) and item["ref_img"].shape == (3, 64, 64)
        assert 0 < float((item["mask"] > 0).float().mean()) < 1
        tr, va = dataloader.get_reference_dataloader(tmp_path / "src", tmp_path / "ref", tmp_path / "mask",
                                                     tmp_path / "identity.txt", 2, num_workers=0, img_scale=0.25)
        assert len(tr.dataset) + len(va.dataset) == 6
    finally:
        sys.path.remove(str(R.REF))
        sys.modules.pop("dataloader", None)


def test_msssim_shim_has_the_classes_the_scripts_import():
    from face_mask_inpaint_b200.patch import _ensure_msssim_shim
    _ensure_msssim_shim()
    from pytorch_msssim import MS_SSIM, SSIM
    a = torch.rand(2, 3, 192, 192)
    b = (a + 0.1 * torch.randn_like(a)).clamp(0, 1)
    s, m = SSIM(data_range=1, size_average=True, channel=3), MS_SSIM(data_range=1, size_average=True, channel=3)
    assert abs(float(s(a, a)) - 1) < 1e-6 and abs(float(m(a, a)) - 1) < 1e-6
    assert 0 < float(s(a, b)) < 1 and 0 < float(m(a, b)) < 1


def test_offline_stubs_build_loss_networks_without_download():
    from face_mask_inpaint_b200.offline import stub_pretrained
    stub_pretrained()
    import torchvision
    v1, v2 = torchvision.models.vgg16(pretrained=True), torchvision.models.vgg16(pretrained=True)
    assert torch.equal(v1.features[0].weight, v2.features[0].weight)      # seeded: every rank builds the same loss network
    sd = torch.hub.load_state_dict_from_url("https://raw.githubusercontent.com/richzhang/PerceptualSimilarity/master/lpips/"
                                            "weights/v0.1/alex.pth")
    assert sd["lin0.model.1.weight"].shape == (1, 64, 1, 1) and float(sd["lin4.model.1.weight"].min()) >= 0
