"""Where does the cold (L2-flushed) step lose its time?  Times one CUDA-graph replay of the fused step
  cold      : after the 512 MiB L2 flush (the protocol bench.py reports)
  data-warm : flush, then the small state buffers of the global part are read (torch .sum()) before the timed replay,
              so only the kernels' CODE (and y) is cold
  warm      : no flush, replays back to back
usage: python tools/coldwarm.py [--workload cfg2]"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='cfg2')
    ap.add_argument('--steps', type=int, default=50)
    a = ap.parse_args()
    args = argparse.Namespace(allreduce='auto', warmup=5, steps=a.steps)
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    w = bench.WORKLOADS[a.workload]
    job = bench.build_job(args, w, 0, 1, dev, None, 'strong')
    fs, one_step = job['fs'], job['one_step']
    for t in range(6):
        one_step(t)
    flush = torch.empty(bench.L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    eng = fs.eng
    state = [fs.theta, fs.adam_m, fs.adam_v, fs.grad, fs.stash, fs.packed, fs.globals, fs.zmask, fs.scalars,
             fs.scalars_in]
    sink = torch.zeros(1, dtype=torch.float64, device=dev)

    def run(mode):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
        for t in range(a.steps):
            if mode != 'warm':
                flush.fill_(t & 0xff)
            if mode == 'data-warm':
                for b in state:
                    sink.add_(b.view(torch.uint8).sum(dtype=torch.float64) if b.dtype != torch.float64 else b.sum())
            ev[t][0].record()
            fs.graph_step()
            ev[t][1].record()
        torch.cuda.synchronize()
        ts = sorted(x.elapsed_time(y) for x, y in ev)
        return sum(ts) / len(ts), ts[len(ts) // 2]

    for mode in ('cold', 'data-warm', 'warm', 'cold'):
        mean, med = run(mode)
        print('%-10s mean %.4f ms  median %.4f ms' % (mode, mean, med), flush=True)


if __name__ == '__main__':
    main()
