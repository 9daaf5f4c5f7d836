"""TEST INFRASTRUCTURE (oracle): plain restatement of the per-image offset loop of the reference's training / test
loops (cad_recognition/train.py:238-258, :340-360): for every key of `slices` containing 'edge' the image's node offset
`slices['pos'][i]` is added to its rows, for 'bbox_idx' the image's proposal offset `slices['labels'][i]`.  Used by the
tests to pin csrc/slicing.cu::k_batch_offsets; never imported by the product path."""
import numpy as np


def apply_offsets(edge, bbox_idx, tab):
    """edge [E,2] int64, bbox_idx [N] int64, tab [4][G+1] = edge | pos | bbox_idx | labels slices.  Returns new arrays."""
    edge = np.array(edge, dtype=np.int64, copy=True)
    bbox_idx = np.array(bbox_idx, dtype=np.int64, copy=True)
    tab = np.asarray(tab, dtype=np.int64)
    G = tab.shape[1] - 1
    for i in range(G):
        edge[tab[0, i]:tab[0, i + 1]] += tab[1, i]
        bbox_idx[tab[2, i]:tab[2, i + 1]] += tab[3, i]
    return edge, bbox_idx
