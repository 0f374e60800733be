"""Oracle-backed stand-in for pipeline.CudaEngine so that the multi-rank HOST logic of TVCScorer
(partitioning, candidate all-to-all, row fetch, histogram all-reduce) can run on CPU with gloo.
Test infrastructure only."""
import numpy as np
import torch

from oracle import tvc_oracle as O


class _G:
    def __init__(self, rows, offset):
        self.rows = np.ascontiguousarray(rows, dtype=np.float32)
        self.offset = int(offset)


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


class OracleEngine:
    device = torch.device("cpu")

    def make_gallery(self, rows, offset, normalize=False):
        rows = _np(rows)
        return _G(O.l2_normalize(rows) if normalize else rows, offset)

    def wrap_rows(self, rows, offset=0):
        return _G(_np(rows), offset)

    def search(self, gallery, q, k, threshold=-np.inf):
        s, i = O.search(_np(q), gallery.rows, k, threshold=threshold, index_offset=gallery.offset)
        return torch.from_numpy(s), torch.from_numpy(i)

    def merge(self, sims, idx, k):
        s, i = O.merge_topk(_np(sims), _np(idx), k)
        return torch.from_numpy(s), torch.from_numpy(i)

    def get_rows(self, gallery, local_idx):
        return torch.from_numpy(gallery.rows[_np(local_idx)])

    def consistency(self, params, img, txt, var, ret_gallery, ret_idx, gen, g_cnt, gen_gallery, gen_idx):
        p = params.as_dict()
        scores, flags, _ = O.consistency_emb(
            _np(img), _np(txt), _np(var),
            ret_rows=ret_gallery.rows if ret_gallery is not None and ret_idx is not None else None,
            ret_idx=_np(ret_idx) if ret_idx is not None else None,
            gen=_np(gen) if gen is not None else None, g_cnt=_np(g_cnt) if g_cnt is not None else None,
            gen_rows=gen_gallery.rows if gen_gallery is not None and gen_idx is not None else None,
            gen_idx=_np(gen_idx) if gen_idx is not None else None, params=p,
            ret_offset=ret_gallery.offset if ret_gallery is not None else 0,
            gen_offset=gen_gallery.offset if gen_gallery is not None else 0)
        return torch.from_numpy(scores.astype(np.float32)), torch.from_numpy(flags)

    def k_occurrence(self, idx, n_bins, counts):
        counts += torch.from_numpy(O.k_occurrence(_np(idx), n_bins))
        return counts
