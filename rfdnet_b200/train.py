"""Data-parallel training step of the joint ISCNet phase (BASELINE config 5).

One process per GPU, batch sharded by scene, per-rank BatchNorm statistics (= the reference's per-replica DataParallel
BN), and ONE exchange step: the gradient all-reduce -- bucketed, launched from autograd hooks while backward is still
running, averaged inside NCCL.  It replaces the reference's single-process nn.DataParallel (net_utils/utils.py:238:
per-step parameter broadcast + output gather + gradient reduce onto GPU 0, loss on GPU 0 training.py:71-74).

  GradBuckets     flat per-bucket gradient buffers (p.grad are views: no flatten / copy-back), one asynchronous
                  all-reduce per bucket as soon as its last gradient has been accumulated
  JointTrainStep  detection (backbone + voting + proposal, train mode) -> SkipPropagation (STN_Group + PointSeg +
                  ResnetPointnet) -> ONet.compute_loss (Encoder_Latent + batch-statistics CBN decoder, KL + BCE),
                  network.py:313-386 with the label matching replaced by fixed synthetic assignments

The point-cloud operators and their gradients underneath (FPS, ball query, group / gather / interpolate and the
scatter-add backward kernels) are librfdnet_b200's; the dense layers train on PyTorch's library GEMMs.
The reference's DetectionLoss / get_proposal_id / NMS (models/loss.py, network.py:182-303) are outside the hot path
(SURVEY.md section 2): a surrogate detection loss over the same head outputs and label tensors is used, so that every
parameter and every backward kernel takes part.
"""
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import completion, detection


class GradBuckets:
    """Bucketed, overlapped gradient all-reduce.

    Parameters are packed -- in REVERSE registration order, the order in which backward produces their gradients --
    into flat fp32 buckets of ~`bucket_bytes`; every p.grad is a view into its bucket, so no flatten or copy-back pass
    exists.  A post-accumulate hook counts a bucket's ready gradients and issues its all-reduce (async, on NCCL's stream)
    the moment the last one arrives; `finish()` waits for the outstanding ones after backward."""

    def __init__(self, params, bucket_bytes=8 << 20, average=True):
        self.params = [p for p in params if p.requires_grad]
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.average = average
        self.buckets = []          # dict(buf, params, pending)
        cur, cur_n = [], 0
        for p in reversed(self.params):
            cur.append(p)
            cur_n += p.numel()
            if cur_n * 4 >= bucket_bytes:
                self._close(cur, cur_n)
                cur, cur_n = [], 0
        if cur:
            self._close(cur, cur_n)
        self._works = []
        self.nbytes = sum(b["buf"].numel() * 4 for b in self.buckets)
        self._of = {}
        for bi, b in enumerate(self.buckets):
            for p in b["params"]:
                self._of[p] = bi
                p.register_post_accumulate_grad_hook(self._hook)

    def _close(self, plist, n):
        dev = plist[0].device
        buf = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in plist:
            p.grad = buf[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.buckets.append({"buf": buf, "params": plist, "pending": len(plist), "sent": False})

    def zero(self):
        """start of a step: zero the buckets (one memset each) and re-arm the counters"""
        for b in self.buckets:
            b["buf"].zero_()
            b["pending"], b["sent"] = len(b["params"]), False
            off = 0
            for p in b["params"]:
                if p.grad is None or p.grad.data_ptr() != b["buf"].data_ptr() + 4 * off:
                    p.grad = b["buf"][off:off + p.numel()].view_as(p)   # someone replaced it (zero_grad(set_to_none))
                off += p.numel()
        self._works = []

    def _send(self, b):
        b["sent"] = True
        if self.world == 1:
            return
        if self.average and dist.get_backend() == "nccl":
            self._works.append(dist.all_reduce(b["buf"], op=dist.ReduceOp.AVG, async_op=True))
        else:
            self._works.append(dist.all_reduce(b["buf"], op=dist.ReduceOp.SUM, async_op=True))

    def _hook(self, p):
        b = self.buckets[self._of[p]]
        b["pending"] -= 1
        if b["pending"] == 0 and not b["sent"]:
            self._send(b)

    def finish(self):
        """after backward: send the buckets that never filled up (parameters without a gradient this step), wait"""
        for b in self.buckets:
            if not b["sent"]:
                self._send(b)
        for w in self._works:
            w.wait()
        if self.world > 1 and self.average and dist.get_backend() != "nccl":
            for b in self.buckets:
                b["buf"].div_(self.world)
        self._works = []

    def allreduce_only(self):
        """the exchange step on its own (for timing): every bucket, back to back, synchronously"""
        for b in self.buckets:
            if self.world > 1:
                dist.all_reduce(b["buf"], op=dist.ReduceOp.SUM)


def synthetic_labels(batch, num_points, boxes_per_scene, points_per_object, device, seed=0):
    """Label tensors with the dataloader's keys / dtypes / shapes (dataloader.py:145-176), random."""
    g = torch.Generator().manual_seed(seed)
    K, T = boxes_per_scene, points_per_object
    lab = {
        "vote_label": torch.randn(batch, num_points, 9, generator=g) * 0.2,
        "vote_label_mask": (torch.rand(batch, num_points, generator=g) > 0.5).long(),
        "objectness_label": (torch.rand(batch, 256, generator=g) > 0.5).long(),
        "sem_cls_label": torch.randint(0, 8, (batch, 256), generator=g),
        "center_label": torch.rand(batch, 256, 3, generator=g) * 4 - 2,
        "point_instance_labels": torch.randint(0, K + 1, (batch, num_points), generator=g).float(),
        "proposal_instance_labels": torch.arange(1, K + 1).float().expand(batch, K).contiguous(),
        "object_points": torch.rand(batch, K, T, 3, generator=g) - 0.5,
        "object_points_occ": (torch.rand(batch, K, T, generator=g) > 0.6).float(),
    }
    return {k: v.to(device) for k, v in lab.items()}


class JointTrainStep(nn.Module):
    """ISCNet's joint phase: detection + SkipPropagation + ONet, one loss (network.py:313-386)."""

    def __init__(self, boxes_per_scene=10, input_feature_dim=1):
        super().__init__()
        self.detection = detection.DetectionHotPath(input_feature_dim, 256)
        self.skip_propagation = completion.SkipPropagation(input_feature_dim=input_feature_dim, c_dim=512, hidden_dim=512)
        self.completion = completion.ONet(z_dim=32, c_dim=512)
        self.K = boxes_per_scene

    def forward(self, point_clouds, lab):
        ep, prop_feat = self.detection(point_clouds, export_proposal_feature=True)
        # ---- surrogate detection loss over every head (see the module docstring)
        seed_inds = ep["seed_inds"].long()
        gt_votes = torch.gather(lab["vote_label"][..., :3], 1, seed_inds.unsqueeze(-1).expand(-1, -1, 3))
        vmask = torch.gather(lab["vote_label_mask"], 1, seed_inds).float()
        vote_loss = ((ep["vote_xyz"] - ep["seed_xyz"] - gt_votes).abs().sum(-1) * vmask).sum() / (vmask.sum() + 1e-6)
        obj_loss = F.cross_entropy(ep["objectness_scores"].transpose(2, 1), lab["objectness_label"])
        sem_loss = F.cross_entropy(ep["sem_cls_scores"].transpose(2, 1), lab["sem_cls_label"])
        box_loss = ((ep["center"] - lab["center_label"]).pow(2).mean() + ep["heading_scores"].pow(2).mean()
                    + ep["heading_residuals_normalized"].pow(2).mean() + ep["size_scores"].pow(2).mean()
                    + ep["size_residuals_normalized"].pow(2).mean())
        det_loss = vote_loss + 0.5 * obj_loss + 0.1 * sem_loss + box_loss
        # ---- completion on K kept proposals per scene (the reference keeps those matched to ground-truth boxes)
        K = self.K
        box_xyz = ep["center"][:, :K].detach().contiguous()
        heading = torch.argmax(ep["heading_scores"][:, :K].detach(), -1).float() * (2 * 3.14159265 / 12)
        codes, mask_loss = self.skip_propagation(box_xyz, heading, prop_feat[:, :, :K].contiguous(), point_clouds,
                                                 lab["point_instance_labels"], lab["proposal_instance_labels"])
        B = point_clouds.shape[0]
        codes = codes.transpose(1, 2).contiguous().view(B * K, -1)
        comp_loss, _ = self.completion.compute_loss(codes, lab["object_points"].view(B * K, -1, 3),
                                                    lab["object_points_occ"].view(B * K, -1), None)
        return det_loss + mask_loss + 0.005 * comp_loss, {"det": det_loss.detach(), "mask": mask_loss.detach(),
                                                          "completion": comp_loss.detach()}


class Trainer:
    """model + Adam + bucketed all-reduce; `step()` = zero, forward, backward (all-reduce overlapped), optimizer."""

    def __init__(self, model, lr=1e-3, bucket_bytes=8 << 20):
        self.model = model
        self.buckets = GradBuckets(list(model.parameters()), bucket_bytes=bucket_bytes)
        self.opt = torch.optim.Adam(model.parameters(), lr=lr)
        self.ev = None

    def step(self, point_clouds, lab, record=False):
        self.buckets.zero()
        loss, parts = self.model(point_clouds, lab)
        loss.backward()
        if record:   # device time at which backward's last kernel was enqueued vs. all-reduce completion
            e_bwd = torch.cuda.Event(enable_timing=True)
            e_bwd.record()
        self.buckets.finish()
        if record:
            e_comm = torch.cuda.Event(enable_timing=True)
            e_comm.record()
            self.ev = (e_bwd, e_comm)
        self.opt.step()
        return loss.detach(), parts
