"""Device-side time-feature ETL: time bins, hours and weekdays of a batch computed by
`cast_time_features` from raw int64 timestamps instead of per-record Python (`reference util.py:24-43, 73-120`,
applied per position by `sampler.py:61-72` / `util.py:276-289`)."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .data import time_bin_edges


class TimeFeaturizer:
    """Holds the bin-edge table of one dataset configuration on the device."""

    def __init__(self, device, bin_in_hours=48, max_bins=200, log_scale=False, min_ts=None, max_ts=None, lib=None):
        self.lib = lib if lib is not None else _lib.load_library()
        self.device = torch.device(device)
        edges = time_bin_edges(bin_in_hours, max_bins, log_scale, min_ts, max_ts)
        self.edges = torch.from_numpy(np.ascontiguousarray(edges)).to(self.device)

    def into(self, ts_dev, ids_dev, out3, stream=None):
        """device tensors in, device tensors out, nothing allocated: ts_dev [B,T] int64, ids_dev [B*T] int32, out3
        [3, B*T] int32 (bins | hours | weekdays — the engine's context-id buffer)"""
        B, T = ts_dev.shape
        if stream is None and self.device.type == "cuda":
            stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self.lib.cast_time_features(ts_dev.data_ptr(), None, ids_dev.data_ptr(), B, T, self.edges.data_ptr(),
                                         int(self.edges.numel()), out3[0].data_ptr(), out3[1].data_ptr(),
                                         out3[2].data_ptr(), stream)
        _lib.check(self.lib, rc, "cast_time_features")

    def __call__(self, ts, ids, ref=None, stream=None):
        """ts [B,T] int64 seconds, ids [B,T] int32 (0 = padding), ref [B] int64 or None (= ts[:, -1]).
        Returns int32 device tensors (bins, hours, days), each [B,T]."""
        ts = torch.as_tensor(ts, dtype=torch.int64).to(self.device).contiguous()
        ids = torch.as_tensor(ids, dtype=torch.int32).to(self.device).contiguous()
        B, T = ts.shape
        out = torch.empty(3, B, T, dtype=torch.int32, device=self.device)
        r = None if ref is None else torch.as_tensor(ref, dtype=torch.int64).to(self.device).contiguous()
        if stream is None and self.device.type == "cuda":
            stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self.lib.cast_time_features(ts.data_ptr(), None if r is None else r.data_ptr(), ids.data_ptr(), B, T,
                                         self.edges.data_ptr(), int(self.edges.numel()), out[0].data_ptr(),
                                         out[1].data_ptr(), out[2].data_ptr(), stream)
        _lib.check(self.lib, rc, "cast_time_features")
        return out[0], out[1], out[2]
