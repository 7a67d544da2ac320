"""Public entry point of the hot path: scenes in, proposals + occupancy logits out.

One call = one pass over a batch of scenes of the path BASELINE.json's metric is quoted on:
  80k-point clouds -> PointNet++ backbone -> votes -> 256 proposals           (ISCNet.forward, network.py:313-330)
  -> ONet decoder queried on the dense 32^3 lattice for every proposal         (generator.py:91-97,123-143)
The per-proposal shape code `c` (512-d) is produced in the reference by SkipPropagation, which is outside this
path; callers pass it in (the benchmark uses seeded N(0,1) codes).
"""
import math

import torch
import torch.nn as nn

from . import completion, detection, generator, onet


class GraphedDetection:
    """The whole detection pass (≈45 launches: FPS, prefix proofs, ball queries, nine tcgen05 MLP chains, 3-NN
    interpolation, head decoding) as ONE CUDA graph per input shape: captured after two eager
    warm-up calls (they build the folded / packed weights and size the library's workspaces), replayed on a static input
    buffer.  Every device-side decision of the pass (the FPS prefix proof) is a flag read by the kernels, so the captured
    graph is valid for any input of that shape.  The returned tensors are the graph's static outputs: they are overwritten
    by the next call with the same shape."""

    def __init__(self, detection_module):
        self.det = detection_module
        self.cache = {}

    @torch.no_grad()
    def __call__(self, point_clouds):
        key = (tuple(point_clouds.shape), str(point_clouds.device))
        ent = self.cache.get(key)
        if ent is None:
            static_in = point_clouds.clone()
            cur = torch.cuda.current_stream(point_clouds.device)
            side = torch.cuda.Stream(point_clouds.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(2):
                    self.det(static_in)
            cur.wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            # captured on the stream the warm-up ran on: the library's ball-query workspace is per (device, stream)
            with torch.cuda.graph(graph, stream=side):
                out = self.det(static_in)
            ent = self.cache[key] = (graph, static_in, out)
        graph, static_in, out = ent
        static_in.copy_(point_clouds, non_blocking=True)
        graph.replay()
        return out


class SceneHotPath(nn.Module):
    def __init__(self, input_feature_dim=1, num_proposal=256, z_dim=32, c_dim=512, resolution=32, box_size=1.1,
                 precision='fp16', backbone_precision='x3', head_precision='x3', graph_detection=False):
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
        # graph_detection: replay the detection pass as one CUDA graph (inference, fixed input shape)
        self._graphed = GraphedDetection(self.detection) if graph_detection else None

    def detect(self, point_clouds):
        if self._graphed is not None and not self.training and not torch.is_grad_enabled():
            return self._graphed(point_clouds)
        return self.detection(point_clouds)

    def grid(self, device):
        if self._grid is None or self._grid.device != torch.device(device):
            self._grid = onet.make_3d_grid(self.resolution, self.box_size, device)
        return self._grid

    @torch.no_grad()
    def forward(self, point_clouds, shape_codes, z=None, logits_out=None):
        """point_clouds (B,N,3+F) f32 cuda; shape_codes (B*K,c_dim); returns (end_points, logits (B*K, R^3))."""
        end_points, _ = self.detect(point_clouds)
        nobj = shape_codes.shape[0]
        if z is None:
            z = torch.zeros((nobj, self.z_dim), dtype=torch.float32, device=shape_codes.device)  # prior mean at test time
        logits = self.decoder.decode(self.grid(point_clouds.device), z, shape_codes)
        if logits_out is not None:
            logits_out.copy_(logits, non_blocking=True)
        return end_points, logits

    @torch.no_grad()
    def run_host(self, pc_host, codes_host, logits_host, device, chunks=4, result="logits", mesh_capacity=(24576, 49152)):
        """End-to-end call with HOST (pinned) buffers: H2D of the clouds and shape codes, the pass, D2H of the result and
        of the proposal scores.  The decoder runs in `chunks` object chunks.
          result="logits": ALL logits come back (logits_host (B*K, R^3) pinned); the D2H of a finished chunk overlaps the
                           decoding of the next one on a second stream.
          result="mesh":   what Generator3D hands to its caller (generator.py:145-168): every chunk's logits go through
                           rfd_extract_mesh on the device into shared vertex / triangle pools; only the used part of the
                           pools is copied, chunk by chunk, under the decoding of the following chunks.
                           `mesh_capacity` = (vertices, triangles) reserved per object in the device pools.  The meshes
                           are left in self.last_meshes = (vertices, triangles, ranges) numpy views of pinned buffers.
          result="bits":   occupancy masks only (rfd_occupancy_bits, 1 bit per lattice point) in self.last_bits.
        Returns bytes moved (h2d, d2h); the caller synchronises the device."""
        main = torch.cuda.current_stream(device)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device)
        copy = self._copy_stream
        pc = pc_host.to(device, non_blocking=True)
        codes = codes_host.to(device, non_blocking=True)
        ep, _ = self.detect(pc)
        scores = ep['objectness_scores'].to('cpu', non_blocking=True)
        nobj = codes.shape[0]
        grid = self.grid(device)
        z = torch.zeros((nobj, self.z_dim), dtype=torch.float32, device=device)
        step = (nobj + chunks - 1) // chunks
        h2d = pc_host.numel() * 4 + codes_host.numel() * 4
        if result == "mesh":
            return h2d, self._run_mesh_result(main, copy, grid, z, codes, nobj, step, mesh_capacity, device) + scores.numel() * 4
        if result == "bits":
            # occupancy only (voxel IoU, external/common.py:7-35): 1 bit per lattice point instead of 32
            words = (self.resolution ** 3 + 31) // 32
            if getattr(self, "_bits_host", None) is None or self._bits_host.shape != (nobj, words):
                self._bits_host = torch.empty((nobj, words), dtype=torch.int32).pin_memory()
            for lo in range(0, nobj, step):
                hi = min(nobj, lo + step)
                lg = self.decoder.decode(grid, z[lo:hi], codes[lo:hi].contiguous())
                bits, _ = onet.occupancy_bits(lg, 0.0)
                self._bits_host[lo:hi].copy_(bits, non_blocking=True)
            self.last_bits = self._bits_host
            return h2d, nobj * words * 4 + scores.numel() * 4
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
        d2h = logits_host.numel() * 4 + scores.numel() * 4
        return h2d, d2h

    def _run_mesh_result(self, main, copy, grid, z, codes, nobj, step, capacity, device):
        """decode + rfd_extract_mesh per object chunk into shared pools; the kernels of a chunk run back to back on the
        main stream, so chunk i owns the pool range between the fill levels after chunks i-1 and i.  The host learns a
        chunk's fill level from a 24-byte snapshot, then copies exactly that range on the copy stream while the next
        chunks are still being decoded."""
        mp = self._mesh_pools(nobj, capacity, device)
        mp["totals"].zero_()
        events, bounds = [], []
        for ci, lo in enumerate(range(0, nobj, step)):
            hi = min(nobj, lo + step)
            lg = self.decoder.decode(grid, z[lo:hi], codes[lo:hi].contiguous())
            generator.extract_meshes(lg, self.resolution, 0.5, self.box_size - 1.0, pools=(mp["v"], mp["t"]),
                                     ranges=mp["ranges"][lo:hi], totals=mp["totals"])
            mp["h_snap"][ci].copy_(mp["totals"], non_blocking=True)
            mp["h_ranges"][lo:hi].copy_(mp["ranges"][lo:hi], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(main)
            events.append(ev)
            bounds.append((lo, hi))
        pv = pt = 0
        grow = False
        for ci, ev in enumerate(events):
            ev.synchronize()
            nv, nt, nofit = (int(x) for x in mp["h_snap"][ci])
            if nofit:
                raise RuntimeError(f"run_host: {nofit} object(s) did not fit in the mesh pools ({nv} vertices / {nt} "
                                   f"triangles reserved so far for {nobj} objects); raise mesh_capacity")
            grow = grow or nv > mp["h_v"].shape[0] or nt > mp["h_t"].shape[0]
            if not grow:
                with torch.cuda.stream(copy):
                    mp["h_v"][pv:nv].copy_(mp["v"][pv:nv], non_blocking=True)
                    mp["h_t"][pt:nt].copy_(mp["t"][pt:nt], non_blocking=True)
            pv, pt = nv, nt
        if grow:   # pinned host pools too small (first steps): regrow with headroom, copy everything once
            copy.synchronize()
            mp["h_v"] = torch.empty((pv + pv // 2, 3), dtype=torch.float32).pin_memory()
            mp["h_t"] = torch.empty((pt + pt // 2, 3), dtype=torch.int32).pin_memory()
            mp["h_v"][:pv].copy_(mp["v"][:pv], non_blocking=True)
            mp["h_t"][:pt].copy_(mp["t"][:pt], non_blocking=True)
        main.wait_stream(copy)
        self.last_meshes = (mp["h_v"][:pv].numpy(), mp["h_t"][:pt].numpy(), mp["h_ranges"][:nobj].numpy())
        return pv * 12 + pt * 12 + nobj * 16 + 24 * len(events)

    def _mesh_pools(self, nobj, capacity, device):
        key = (nobj, tuple(capacity), str(device))
        mp = getattr(self, "_mesh_pool_cache", None)
        if mp is None or mp["key"] != key:
            cv, ct = nobj * capacity[0], nobj * capacity[1]
            mp = {"key": key,
                  "v": torch.empty((cv, 3), dtype=torch.float32, device=device),
                  "t": torch.empty((ct, 3), dtype=torch.int32, device=device),
                  "ranges": torch.empty((nobj, 4), dtype=torch.int32, device=device),
                  "totals": torch.zeros((3,), dtype=torch.int64, device=device),
                  "h_v": torch.empty((0, 3), dtype=torch.float32),
                  "h_t": torch.empty((0, 3), dtype=torch.int32),
                  "h_ranges": torch.empty((nobj, 4), dtype=torch.int32).pin_memory(),
                  "h_snap": torch.empty((64, 3), dtype=torch.int64).pin_memory()}
            self._mesh_pool_cache = mp
        return mp


class SceneGeneration(nn.Module):
    """The device part of ISCNet.generate (models/iscnet/modules/network.py:56-153), end to end on this library:

        detection (backbone -> votes -> 256 proposals, proposal features exported)          network.py:66-83
        -> box centres + heading angles of the kept proposals                                network.py:109-119
        -> SkipPropagation.generate: per-proposal shape codes                                network.py:124, skip_propagation.py:84-129
        -> ONet decoder on the dense R^3 lattice, z = prior mean                             generator.py:64-97, occupancy_net.py:147-175
        -> meshes (pad, marching cubes, vertex transform)                                    generator.py:145-168

    What the reference does in between on the HOST -- parse_predictions, 3D NMS and the ground-truth matching that pick
    BATCH_PROPOSAL_IDs (network.py:84-101) -- is outside the hot path: callers pass the proposal ids to
    keep (B, K) or get all proposals.  Sub-module names follow ISCNet's (`skip_propagation`, `completion.decoder`), the
    detection part sits under `detection.*` as in SceneHotPath."""

    def __init__(self, input_feature_dim=1, num_proposal=256, z_dim=32, c_dim=512, resolution=32, padding=0.1,
                 num_heading_bin=12, precision='fp16'):
        super().__init__()
        self.detection = detection.DetectionHotPath(input_feature_dim, num_proposal, num_heading_bin=num_heading_bin)
        self.skip_propagation = completion.SkipPropagation(input_feature_dim=input_feature_dim, c_dim=c_dim, hidden_dim=512)
        self.completion = completion.ONet(z_dim=z_dim, c_dim=c_dim, precision=precision)
        self.num_heading_bin, self.resolution, self.padding, self.z_dim = num_heading_bin, resolution, padding, z_dim
        self._grid = None

    @staticmethod
    def heading_angles(end_points, num_heading_bin):
        """scannet_config.py:55-63 (class2angle_cuda) on the arg-max heading bin + its residual (network.py:112-117)"""
        cls = torch.argmax(end_points['heading_scores'], -1)
        res = end_points['heading_residuals_normalized'] * (math.pi / num_heading_bin)
        res = torch.gather(res, 2, cls.unsqueeze(-1)).squeeze(2)
        angle = cls.float() * (2 * math.pi / num_heading_bin) + res
        return angle - 2 * math.pi * (angle > math.pi).float()

    @torch.no_grad()
    def forward(self, point_clouds, proposal_ids=None, meshes=True):
        """point_clouds (B,N,3+F) f32 cuda; proposal_ids (B,K) int64 or None (= all) ->
        dict(end_points, codes (B*K, c_dim), logits (B*K, R^3), meshes: generator.MeshBatch | None)"""
        ep, prop_feat = self.detection(point_clouds, export_proposal_feature=True)
        centers, angles = ep['center'], self.heading_angles(ep, self.num_heading_bin)
        if proposal_ids is not None:
            ids = proposal_ids.long()
            prop_feat = torch.gather(prop_feat, 2, ids.unsqueeze(1).expand(-1, prop_feat.shape[1], -1))
            centers = torch.gather(centers, 1, ids.unsqueeze(-1).expand(-1, -1, 3))
            angles = torch.gather(angles, 1, ids)
        codes = self.skip_propagation.generate(centers.contiguous(), angles.contiguous(), prop_feat.contiguous(), point_clouds)
        B, C, K = codes.shape
        codes = codes.transpose(1, 2).contiguous().view(B * K, C)
        dev = point_clouds.device
        if self._grid is None or self._grid.device != dev:
            self._grid = onet.make_3d_grid(self.resolution, 1 + self.padding, dev)
        z = torch.zeros((B * K, self.z_dim), dtype=torch.float32, device=dev)            # prior mean (generator.py:79)
        logits = self.completion.decoder.decode(self._grid, z, codes)
        mb = generator.extract_meshes(logits, self.resolution, 0.5, self.padding, vertices_per_object=24576,
                                      triangles_per_object=49152) if meshes else None
        return {"end_points": ep, "codes": codes, "logits": logits, "meshes": mb}
