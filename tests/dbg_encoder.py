import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from face_mask_inpaint_b200.modules.picnet import build_picnet_ref
from golden_util import fill_by_name
m = fill_by_name(build_picnet_ref()).eval().cuda()
img = torch.rand(2, 3, 64, 96).cuda()
with torch.no_grad():
    (mu, std), f = m.ref_encoder(img)
torch.cuda.synchronize()
print(mu.shape, f.shape, float(f.abs().mean()))
