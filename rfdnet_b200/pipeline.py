"""Public entry point of the hot path: scenes in, proposals + occupancy logits out.

One call = one pass over a batch of scenes of the path BASELINE.json's metric is quoted on:
  80k-point clouds -> PointNet++ backbone -> votes -> 256 proposals           (ISCNet.forward, network.py:313-330)
  -> ONet decoder queried on the dense 32^3 lattice for every proposal         (generator.py:91-97,123-143)
The per-proposal shape code `c` (512-d) is produced in the reference by SkipPropagation, which is outside this
path; callers pass it in (the benchmark uses seeded N(0,1) codes).
"""
import torch
import torch.nn as nn

from . import detection, onet


class SceneHotPath(nn.Module):
    def __init__(self, input_feature_dim=1, num_proposal=256, z_dim=32, c_dim=512, resolution=32, box_size=1.1,
                 precision='fp16', backbone_precision='x3', head_precision='x3'):
        """precision: ONet decoder ('fp16' | 'fp16x3' | 'bf16' tcgen05 modes, 'fp32' CUDA cores).
        backbone_precision: the MLPs of SA1-4 / FP1-2 (BASELINE config 2 is fp32: 'x3' = split-fp16 tensor cores with
        fp32-grade results; 'fp16' / 'bf16' single-MMA tensor-core modes; 'cuda' = fp32 CUDA-core layer kernel).
        head_precision: voting MLP, vote-aggregation SA layer and proposal head (BASELINE config 3), same choices."""
        super().__init__()
        self.detection = detection.DetectionHotPath(input_feature_dim, num_proposal)
        for name in ("sa1", "sa2", "sa3", "sa4", "fp1", "fp2"):
            getattr(self.detection.backbone, name).precision = backbone_precision
        self.detection.voting.precision = head_precision
        self.detection.detection.precision = head_precision
        self.detection.detection.vote_aggregation.precision = head_precision
        self.decoder = onet.DecoderCBatchNorm(dim=3, z_dim=z_dim, c_dim=c_dim, precision=precision)
        self.num_proposal, self.z_dim, self.c_dim = num_proposal, z_dim, c_dim
        self.resolution, self.box_size = resolution, box_size
        self._grid = None

    def grid(self, device):
        if self._grid is None or self._grid.device != torch.device(device):
            self._grid = onet.make_3d_grid(self.resolution, self.box_size, device)
        return self._grid

    @torch.no_grad()
    def forward(self, point_clouds, shape_codes, z=None, logits_out=None):
        """point_clouds (B,N,3+F) f32 cuda; shape_codes (B*K,c_dim); returns (end_points, logits (B*K, R^3))."""
        end_points, _ = self.detection(point_clouds)
        nobj = shape_codes.shape[0]
        if z is None:
            z = torch.zeros((nobj, self.z_dim), dtype=torch.float32, device=shape_codes.device)  # prior mean at test time
        logits = self.decoder.decode(self.grid(point_clouds.device), z, shape_codes)
        if logits_out is not None:
            logits_out.copy_(logits, non_blocking=True)
        return end_points, logits

    @torch.no_grad()
    def run_host(self, pc_host, codes_host, logits_host, device, chunks=4):
        """End-to-end call with HOST (pinned) buffers: H2D of the clouds and shape codes, the pass, D2H of ALL logits
        and of the proposal scores.  The decoder runs in `chunks` object chunks; the D2H copy of a finished chunk
        overlaps the decoding of the next one on a second stream.  Returns bytes moved (h2d, d2h); the caller
        synchronises the device (both streams are joined before returning control of the buffers)."""
        main = torch.cuda.current_stream(device)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device)
        copy = self._copy_stream
        pc = pc_host.to(device, non_blocking=True)
        codes = codes_host.to(device, non_blocking=True)
        ep, _ = self.detection(pc)
        scores = ep['objectness_scores'].to('cpu', non_blocking=True)
        nobj = codes.shape[0]
        grid = self.grid(device)
        z = torch.zeros((nobj, self.z_dim), dtype=torch.float32, device=device)
        step = (nobj + chunks - 1) // chunks
        keep = []
        for lo in range(0, nobj, step):
            hi = min(nobj, lo + step)
            lg = self.decoder.decode(grid, z[lo:hi], codes[lo:hi].contiguous())
            ev = torch.cuda.Event()
            ev.record(main)
            copy.wait_event(ev)
            with torch.cuda.stream(copy):
                logits_host[lo:hi].copy_(lg, non_blocking=True)
            lg.record_stream(copy)
            keep.append(lg)
        main.wait_stream(copy)
        h2d = pc_host.numel() * 4 + codes_host.numel() * 4
        d2h = logits_host.numel() * 4 + scores.numel() * 4
        return h2d, d2h
