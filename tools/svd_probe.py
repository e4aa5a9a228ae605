"""Dump torch.svd's outputs on the GPU (cuSOLVER batched Jacobi) for random 2x2 covariance matrices, to characterise its
singular-vector sign convention offline.  Writes gpurun_out/svd_probe.pt."""
import torch

g = torch.Generator().manual_seed(0)
n = 200000
# covariances of the region predictor: PSD, both orders of (a, c), both signs of b, some nearly isotropic / degenerate
l = torch.randn(n, 2, 2, generator=g) * torch.rand(n, 1, 1, generator=g)
cov = l @ l.transpose(1, 2) + 1e-4 * torch.eye(2)
cov[:1000, 0, 1] = cov[:1000, 1, 0] = 0.0                      # diagonal matrices
cov[1000:2000] = torch.eye(2) * torch.rand(1000, 1, 1, generator=g)      # isotropic
out = {"cov": cov}
for bs in (n, 640, 32):                                        # the path may depend on the batch size
    c = cov[:bs].cuda()
    u, s, v = torch.svd(c)
    out[f"u_{bs}"], out[f"s_{bs}"], out[f"v_{bs}"] = u.cpu(), s.cpu(), v.cpu()
u, s, v = torch.svd(cov[:20000])
out["u_cpu"], out["s_cpu"], out["v_cpu"] = u, s, v
torch.save(out, "gpurun_out/svd_probe.pt")
print("saved", {k: tuple(t.shape) for k, t in out.items()})
