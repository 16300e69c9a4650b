import copy, os, sys, types, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from face_mask_inpaint_b200.modules.picnet import build_picnet_ref
from golden_util import fill_by_name, mean_z, picnet_inputs
def rel(a, b): return ((a - b).abs().max() / b.abs().max()).item()
base = fill_by_name(build_picnet_ref()).eval()
src, ref, mask = (t.cuda() for t in picnet_inputs(1))
res = {}
for batch in ("1", "0", "0"):
    os.environ["FMI_SN_BATCH"] = batch
    m = copy.deepcopy(base).cuda()
    m.decoder.get_z = types.MethodType(mean_z, m.decoder)
    with torch.no_grad():
        o1 = m(src, ref, mask).clone(); o2 = m(src, ref, mask).clone(); o3 = m(src, ref, mask).clone()
    key = batch if batch not in res else batch + "b"
    res[key] = (o1, o2, o3, {n: p.detach().clone() for n, p in m.named_parameters() if n.endswith("_u") or n.endswith("_v")})
for a, b in (("1", "0"), ("0b", "0")):
    print(a, "vs", b, "images:", [f"{rel(x, y):.2e}" for x, y in zip(res[a][:3], res[b][:3])])
    worst = sorted(((rel(res[a][3][n], res[b][3][n]), n) for n in res[a][3]), reverse=True)[:5]
    print("  worst u/v:", [(f"{e:.2e}", n) for e, n in worst])
