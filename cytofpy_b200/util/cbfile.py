"""Cord-blood text format of reference cytofpy/util/readCB.py:3-40: a header
line `<label> N_1 ... N_I`, a line of marker names, then one whitespace-separated
row per cell with `NA` for missing values."""
import numpy as np
import torch


def readCB(cb_filepath):
    """Returns {'N', 'markers', 'y': [(N_i, J) fp64], 'm': [isnan masks]} (readCB.py:3-28)."""
    with open(cb_filepath, 'r') as f:
        N = [int(tok) for tok in f.readline().split()[1:]]
        markers = f.readline().split()
        body = f.read().replace('NA', 'nan').split()
    mat = np.asarray(body[:sum(N) * len(markers)], dtype=np.float64).reshape(sum(N), len(markers))
    y = [torch.tensor(part) for part in np.split(mat, np.cumsum(N)[:-1], axis=0)]
    return {'N': N, 'markers': markers, 'y': y, 'm': [torch.isnan(t) for t in y]}


def preprocess(data, rm_cells_below=-6.0):
    """Keeps cells whose every marker is above the threshold or missing (readCB.py:30-40);
    mutates `data` like the reference and returns None."""
    for i, yi in enumerate(data['y']):
        keep = ((yi > rm_cells_below) | torch.isnan(yi)).all(1)
        data['N'][i] = int(keep.sum())
        data['y'][i] = yi[keep]
        if 'm' in data:
            data['m'][i] = data['m'][i][keep]
