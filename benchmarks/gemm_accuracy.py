"""Where does the f32 GEMM's error come from?  ours vs cuBLAS SGEMM (fp32) vs cuBLAS 1xTF32, all against fp64."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
import ndarray_b200 as nd  # noqa: E402
from ndarray_b200 import Array  # noqa: E402


def main():
    ctx = nd.default_context()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    for k in (32, 128, 1024, 4096, 8192, 32768):
        m = n = 512
        a = torch.randn(m, k, device=dev)
        b = torch.randn(k, n, device=dev)
        c = torch.zeros(m, n, device=dev)
        A = Array.borrow(a.data_ptr(), (m, k), np.float32, "C", ctx, owner=a)
        B = Array.borrow(b.data_ptr(), (k, n), np.float32, "C", ctx, owner=b)
        Cm = Array.borrow(c.data_ptr(), (m, n), np.float32, "C", ctx, owner=c)
        nd.general_mat_mul(1.0, A, B, 0.0, Cm)
        ctx.sync()
        ref = a.double() @ b.double()
        scale = ref.abs().max()
        torch.backends.cuda.matmul.allow_tf32 = False
        sg = a @ b
        torch.backends.cuda.matmul.allow_tf32 = True
        tf = a @ b
        torch.backends.cuda.matmul.allow_tf32 = False
        e = lambda x: ((x.double() - ref).abs().max() / scale).item()
        bias = ((c.double() - ref).mean() / scale).item()
        print(f"K={k:6d}  ours {e(c):.2e} (mean signed err {bias:+.1e})   cublas fp32 {e(sg):.2e}   cublas 1xTF32 {e(tf):.2e}")
    # a case only the lo terms can get right: every a = 1 + 2^-15 (tf32 cannot hold it), B = identity
    k = 256
    a = torch.full((256, k), 1.0 + 2.0 ** -15, device=dev)
    b = torch.eye(k, device=dev)
    c = torch.zeros(256, k, device=dev)
    A = Array.borrow(a.data_ptr(), (256, k), np.float32, "C", ctx, owner=a)
    B = Array.borrow(b.data_ptr(), (k, k), np.float32, "C", ctx, owner=b)
    Cm = Array.borrow(c.data_ptr(), (256, k), np.float32, "C", ctx, owner=c)
    nd.general_mat_mul(1.0, A, B, 0.0, Cm)
    ctx.sync()
    print("lo-term check: C[0,0] - 1 =", (c[0, 0].double() - 1).item(), "expected", 2.0 ** -15, "exact match:", bool((c == a).all()))


if __name__ == "__main__":
    main()
