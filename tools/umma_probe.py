"""Debug helper: run one tensor-core conv case per process and report error / elapsed time."""
import math, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "constrastive-predictive-coding-audio_b200"))

def run_case(args):
    import torch, torch.nn.functional as F
    import cpc_b200
    b, cin, h, w, cout, kh, kw, ph, pw, top = args
    g = torch.Generator().manual_seed(1)
    x = torch.randn(b, cin, h, w, generator=g); wt = torch.randn(cout, cin, kh, kw, generator=g) / math.sqrt(cin*kh*kw)
    want = F.conv2d(F.pad(x.double(), (0, 0, top, 0)), wt.double(), None, padding=(ph, pw))
    t0 = time.time()
    got = cpc_b200.ops.conv2d(x.cuda(), wt.cuda(), None, (1, 1), (ph, pw), extra_top=top)
    torch.cuda.synchronize()
    err = float((got.cpu().double() - want).norm() / want.norm())
    print("case", args, "rel_err %.3e" % err, "time %.2fs" % (time.time() - t0), flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(tuple(int(v) for v in sys.argv[1].split(",")))
    else:
        cases = ["1,64,9,130,256,2,2,0,0,0", "1,64,9,130,128,2,2,0,0,0", "1,64,9,130,256,1,1,0,0,0", "1,64,9,130,64,2,2,0,0,0",
                 "1,128,9,130,256,2,2,0,0,0", "1,64,9,64,256,2,2,0,0,0", "2,64,9,130,256,2,2,0,0,0", "1,64,9,130,128,1,1,0,0,0"]
        for c in cases:
            t0 = time.time()
            r = subprocess.run([sys.executable, __file__, c], capture_output=True, text=True, env=dict(os.environ, CUDA_LAUNCH_BLOCKING="1"))
            tail = (r.stdout + r.stderr).strip().splitlines()[-1:] 
            print(c, "rc", r.returncode, "%.1fs" % (time.time() - t0), tail, flush=True)
