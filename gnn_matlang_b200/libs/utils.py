"""Drop-in replacement for the reference's ``libs/utils.py::SpectralDesign`` (libs/utils.py:525-626):

    from gnn_matlang_b200.libs.utils import SpectralDesign
    transform = SpectralDesign(nmax=0, recfield=1, dv=5, nfreq=5, adddegree=False, laplacien=True, addadj=False, vmax=None)
    data = transform(data)          # adds edge_index2, edge_attr2, lmax; augments x if adddegree

Same constructor arguments and the same ``__call__(data) -> data`` contract, plus a batched entry point
(``design_batch``) that designs the supports of many graphs in one launch -- one CUDA thread block per graph,
FP64 Jacobi eigensolver in shared memory (the reference loops over graphs in Python with one dense numpy
``eigh`` each).  Graphs beyond the shared-memory envelope of that kernel (more than ``gnnml3_spectral_max_nodes`` nodes:
the single 900-node grid of ``filtering.py:17``) take a dense device path (``_design_large``: cuSOLVER ``eigh`` + dense
products through torch, still no CPU arithmetic) -- outside the north-star's "small symmetric Laplacians", kept so that the
reference's script finds its transform.  The PPGN tensors ``X2`` / ``M`` that the reference also builds when ``nmax > 0``
(:613-624) are not read by GNNML3 and are not produced (``nmax`` is accepted and ignored).
No CPU fallback: a CUDA device is required.
"""
import numpy as np
import torch

from .. import _lib


def get_n_params(model):
    """libs/utils.py:14-21"""
    return sum(p.numel() for p in model.parameters())


class SpectralDesign(object):

    def __init__(self, nmax=0, recfield=1, dv=5, nfreq=5, adddegree=False, laplacien=True, addadj=False, vmax=None):
        self.recfield = recfield        # receptive field. 0: adj, 1: adj+I, r: 2^(r-1)-hop area
        self.dv = dv                    # bandwidth of the Gaussian band filters
        self.nfreq = nfreq              # number of sampled points of the spectrum
        self.adddegree = adddegree      # append the node degree to the node features
        self.laplacien = laplacien      # spectrum of the normalised Laplacian (True) or of the adjacency (False)
        self.addadj = addadj            # add the adjacency as one more edge feature
        self.vmax = vmax                # use the given maximum eigenvalue
        self.nmax = nmax                # PPGN only; ignored here

    @property
    def num_supports(self):
        return self.nfreq + 1 + (1 if self.addadj else 0)

    # ------------------------------------------------------------------------------------------ batched
    def design_batch(self, edge_index, edge_ptr, node_ptr, device=None, global_ids=False, max_entries=None, nmax=None,
                     num_nodes=None):
        """edge_index [2, Etot] int64 with LOCAL node ids, edge_ptr / node_ptr [B+1] (any int dtype, any device).
        Returns a dict of device tensors: edge_index2 [2, E2], edge_attr2 [E2, K], e2_ptr [B+1] (int64),
        lmax [B], degree [Ntot].

        The number of mask entries E2 is data dependent: by default it is read back from the device (one host
        synchronisation per call).  With ``max_entries`` (an upper bound, e.g. ``sum(n_b ** 2)``) nothing is read back:
        the outputs have ``max_entries`` columns / rows, the tail beyond ``e2_ptr[-1]`` holds entries (0, 0) with all-zero
        supports -- exactly neutral for SpectConv / ML3Layer (they add 0 to node 0 and receive zero gradients) -- so the
        call is CUDA-graph capturable and the downstream CSR build needs no size from the device.

        ``node_ptr`` may be a DEVICE tensor when the caller also passes ``nmax`` (largest graph) and ``num_nodes`` (total):
        nothing is uploaded then (uploading a pageable host tensor synchronises the stream: with designs of several batches in
        flight on different streams that upload was the serialisation point -- train.DesignFeeder)."""
        lib = _lib.load()
        if device is None:
            device = edge_index.device if edge_index.is_cuda else torch.device("cuda", torch.cuda.current_device())
        if not torch.cuda.is_available():
            raise RuntimeError("gnn_matlang_b200.SpectralDesign needs a CUDA device (no CPU fallback)")
        if isinstance(node_ptr, torch.Tensor) and node_ptr.is_cuda:
            if nmax is None or num_nodes is None:
                raise ValueError("design_batch: a device node_ptr needs nmax= and num_nodes=")
            B = node_ptr.numel() - 1
            npd = node_ptr.to(dtype=torch.int32).contiguous()
            Ntot_given = int(num_nodes)
            nmax = int(nmax)
        else:
            node_ptr_h = torch.as_tensor(node_ptr).to("cpu", torch.int64)
            B = node_ptr_h.numel() - 1
            nmax = int((node_ptr_h[1:] - node_ptr_h[:-1]).max()) if B > 0 else 1
            npd = node_ptr_h.to(device=device, dtype=torch.int32).contiguous()
            Ntot_given = int(node_ptr_h[-1]) if B > 0 else 0
        ei = torch.as_tensor(edge_index).to(device=device, dtype=torch.int64).contiguous()
        ep = torch.as_tensor(edge_ptr).to(device=device, dtype=torch.int32).contiguous()
        Etot, Ntot = ei.size(1), Ntot_given
        K = self.num_supports
        counts = torch.zeros(B, dtype=torch.int32, device=device)
        with torch.cuda.device(device):
            st = _lib.stream_ptr()
            _lib.check(lib.gnnml3_spectral_count(_lib.ptr(ei), Etot, _lib.ptr(ep), _lib.ptr(npd), B, int(self.recfield),
                                                 max(nmax, 1), _lib.ptr(counts), st), "gnnml3_spectral_count")
            e2_ptr = torch.zeros(B + 1, dtype=torch.int64, device=device)
            torch.cumsum(counts, 0, out=e2_ptr[1:])
            if max_entries is not None:
                E2 = int(max_entries)
                ei2 = torch.zeros(2, E2, dtype=torch.int64, device=device)
                ea2 = torch.zeros(E2, K, dtype=torch.float32, device=device)
            else:
                E2 = int(e2_ptr[-1].item()) if B > 0 else 0
                ei2 = torch.empty(2, E2, dtype=torch.int64, device=device)
                ea2 = torch.empty(E2, K, dtype=torch.float32, device=device)
            lmax = torch.zeros(B, dtype=torch.float32, device=device)
            deg = torch.zeros(Ntot, dtype=torch.float32, device=device)
            _lib.check(lib.gnnml3_spectral_design(
                _lib.ptr(ei), Etot, _lib.ptr(ep), _lib.ptr(npd), B, int(self.recfield), float(self.dv), int(self.nfreq),
                int(bool(self.laplacien)), int(bool(self.addadj)), int(self.vmax is not None),
                float(self.vmax if self.vmax is not None else 0.0), max(nmax, 1), _lib.ptr(e2_ptr), int(bool(global_ids)),
                _lib.ptr(ei2), E2, _lib.ptr(ea2), _lib.ptr(lmax), _lib.ptr(deg), st), "gnnml3_spectral_design")
        return dict(edge_index2=ei2, edge_attr2=ea2, e2_ptr=e2_ptr, counts=counts, lmax=lmax, degree=deg)

    # ------------------------------------------------------------------------------------------ one large graph
    def max_kernel_nodes(self):
        """Largest graph the one-block-per-graph kernel designs (its Jacobi lives in shared memory)."""
        return int(_lib.load().gnnml3_spectral_max_nodes(int(self.nfreq)))

    def _design_large(self, edge_index, n, device=None):
        """One graph of more than ``max_kernel_nodes()`` nodes, dense on the device (libs/utils.py:558-610 restated with
        torch ops: FP64 ``torch.linalg.eigh`` = cuSOLVER for the Laplacian, the reference's float32 ``eigh(A)`` when
        ``laplacien=False``).  Same outputs as ``design_batch`` with B = 1; the mask is binarised after every squaring (same
        pattern as the reference's ``(A + I)^(2^(r-1)) > 0``: all entries are non-negative)."""
        if not torch.cuda.is_available():
            raise RuntimeError("gnn_matlang_b200.SpectralDesign needs a CUDA device (no CPU fallback)")
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        ei = torch.as_tensor(edge_index).reshape(2, -1).to(device=device, dtype=torch.int64)
        f64 = dict(dtype=torch.float64, device=device)
        A = torch.zeros(n, n, dtype=torch.float32, device=device)
        A[ei[0], ei[1]] = 1.0                                                     # :558-560
        deg = A.sum(0)                                                            # column sums (:563, :576)
        eye = torch.eye(n, **f64)
        if self.recfield == 0:                                                    # :566-573
            M = A > 0
        else:
            M = (A.double() + eye) > 0
            for _ in range(1, int(self.recfield)):
                Md = M.double()
                M = (Md @ Md) > 0
        dis = 1.0 / deg.sqrt()
        dis = torch.where(torch.isfinite(dis), dis, torch.zeros_like(dis))        # :578-580
        AD = A * dis[None, :]                                                     # float32 products as in the reference (:582)
        nL = eye - (AD.t() * dis[None, :]).double()
        V, U = torch.linalg.eigh(nL)                                              # :583-584
        V = V.clamp_min(0.0)
        lmax = V.max().to(torch.float32)                                          # :586
        if not self.laplacien:                                                    # :588-589 (float32, unclamped)
            V, U = torch.linalg.eigh(A)
            V, U = V.double(), U.double()
        top = V.max() if self.vmax is None else torch.tensor(float(self.vmax), **f64)          # :592-596
        lo = V.min()
        nf = int(self.nfreq)
        centers = [lo + (top - lo) * (i / (nf - 1)) if nf > 1 else lo for i in range(nf)]       # np.linspace(V.min(), vmax, nfreq)
        r, c = torch.nonzero(M, as_tuple=True)                                    # row-major, as np.where (:608)
        K = self.num_supports
        ea2 = torch.zeros(r.numel(), K, dtype=torch.float32, device=device)
        for i in range(nf):                                                       # :599-600
            S = (U * torch.exp(-(float(self.dv) * (V - centers[i]) ** 2))[None, :]) @ U.t()
            ea2[:, i] = S[r, c].to(torch.float32)
        ea2[:, nf] = (r == c).to(torch.float32)                                   # identity support (:602)
        if self.addadj:
            ea2[:, nf + 1] = A[r, c]                                              # :604-605
        e2_ptr = torch.tensor([0, r.numel()], dtype=torch.int64, device=device)
        return dict(edge_index2=torch.stack([r, c]), edge_attr2=ea2, e2_ptr=e2_ptr,
                    counts=torch.tensor([r.numel()], dtype=torch.int32, device=device), lmax=lmax.reshape(1), degree=deg)

    def design_list(self, graphs, device=None):
        """``graphs``: list of (n, edge_index [2,e]) pairs -> per-graph list of dicts (edge_index2, edge_attr2,
        lmax, degree), all designed in one launch."""
        limit = self.max_kernel_nodes()
        if any(int(n) > limit for n, _ in graphs):
            # graphs beyond the kernel's envelope one by one on the dense path, the rest in one launch
            small = [(i, g) for i, g in enumerate(graphs) if int(g[0]) <= limit]
            res = [None] * len(graphs)
            for (i, _), r in zip(small, self.design_list([g for _, g in small], device=device) if small else []):
                res[i] = r
            for i, (n, e) in enumerate(graphs):
                if int(n) > limit:
                    o = self._design_large(np.asarray(e, dtype=np.int64), int(n), device=device)
                    res[i] = dict(edge_index2=o["edge_index2"], edge_attr2=o["edge_attr2"], lmax=o["lmax"][0], degree=o["degree"])
            return res
        ns = np.array([int(n) for n, _ in graphs], dtype=np.int64)
        es = np.array([int(np.asarray(e).shape[1]) for _, e in graphs], dtype=np.int64)
        node_ptr = np.concatenate([[0], np.cumsum(ns)])
        edge_ptr = np.concatenate([[0], np.cumsum(es)])
        ei = np.concatenate([np.asarray(e, dtype=np.int64).reshape(2, -1) for _, e in graphs], 1) if len(graphs) else np.zeros((2, 0), np.int64)
        out = self.design_batch(torch.from_numpy(ei), torch.from_numpy(edge_ptr), torch.from_numpy(node_ptr), device=device)
        e2 = out["e2_ptr"].cpu().numpy()
        res = []
        for b in range(len(graphs)):
            res.append(dict(edge_index2=out["edge_index2"][:, e2[b]:e2[b + 1]], edge_attr2=out["edge_attr2"][e2[b]:e2[b + 1]],
                            lmax=out["lmax"][b], degree=out["degree"][node_ptr[b]:node_ptr[b + 1]]))
        return res

    # ------------------------------------------------------------------------------------------ per graph (reference API)
    def __call__(self, data):
        n = data.x.shape[0]
        src_dev = data.x.device
        data.x = data.x.type(torch.float32)
        ei = data.edge_index
        if n > self.max_kernel_nodes():
            out = self._design_large(ei, n)
        else:
            out = self.design_batch(ei.reshape(2, -1), torch.tensor([0, ei.reshape(2, -1).shape[1]]), torch.tensor([0, n]))
        if self.adddegree:
            data.x = torch.cat([data.x, out["degree"].to(src_dev).unsqueeze(-1)], 1)
        data.lmax = np.float32(out["lmax"][0].item())
        data.edge_index2 = out["edge_index2"].to(src_dev)
        data.edge_attr2 = out["edge_attr2"].to(src_dev)
        return data
