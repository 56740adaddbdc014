"""Writes profiles/traffic.json from `ncu --set full` captures of the dominant kernel: per workload the DRAM bytes of
ONE launch (dram__bytes_read.sum + dram__bytes_write.sum) and the sha1 of the kernel source the capture was taken
on.  bench.py reports `roofline.traffic` from this file and prints null when the source has changed since.

    python tools/make_traffic.py cfg4=gpurun_out/r2f_mma_kernel.ncu-rep [cfg2=...] --source cytofpy_b200/csrc/elbo_kernel_mma.cuh
"""
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dram_bytes(rep):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    tot = 0.0
    for name in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        i = hdr.index(name)
        scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[units[i]]
        tot += float(vals[i].replace(',', '')) * scale
    return tot, rows[2][hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''


def main():
    args = [a for a in sys.argv[1:] if '=' in a]
    source = sys.argv[sys.argv.index('--source') + 1]
    sha = hashlib.sha1(open(os.path.join(ROOT, source), 'rb').read()).hexdigest()
    path = os.path.join(ROOT, 'profiles', 'traffic.json')
    out = json.load(open(path)) if os.path.exists(path) else {}
    for a in args:
        wl, rep = a.split('=', 1)
        b, kernel = dram_bytes(rep)
        out[wl] = {'dram_bytes_per_launch': b, 'kernel': kernel, 'capture': os.path.basename(rep), 'source': source,
                   'source_sha1': sha}
        print(wl, b, kernel)
    json.dump(out, open(path, 'w'), indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
