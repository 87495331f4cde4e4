import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, fields
import isoext_b200 as iso
n = 512
vals = fields.eval_field(fields.torus(), (n, n, n)).cuda()
g = iso.UniformGrid([n] * 3); g.set_values(vals)
v, f = iso.marching_cubes(g)
b = torch.floor((v[:, 0].double() + 1) / 2 * (n - 1)).long()
c = torch.bincount(b, minlength=n)
print("V", len(v), "max bucket", int(c.max()), "mean nonzero", float(c[c > 0].float().mean()), "n>2048:", int((c > 2048).sum()), "n>1024:", int((c > 1024).sum()), "n>512:", int((c > 512).sum()), "nonzero", int((c > 0).sum()))
print(c[::16].tolist())
