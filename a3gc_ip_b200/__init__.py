"""a3gc_ip_b200 -- B200-native (sm_100a) implementation of the recurrent adaptive-graph-convolution
hot path of trikpachu/A3GC-IP: drop-in torch.nn.Modules over a C-ABI CUDA library.

    from a3gc_ip_b200 import A3GC_net, PoseNet3      # same API as the reference's net_aagc.py
"""
from . import _lib
from . import synthetic
from ._lib import build, lib, LIB_PATH
from .net_aagc import (AAGC, AAGC_LSTM_cell, A3GC_LSTM_cell, AGC_LSTM_cell, G_GRU_cell,
                       AAGC_LSTM, ReverseAAGC_LSTM, BiAAGC_LSTM, A3GC_LSTM, ReverseA3GC_LSTM, BiA3GC_LSTM,
                       AGC_LSTM, ReverseAGC_LSTM, BiAGC_LSTM, G_GRU, ReverseG_GRU, BiG_GRU,
                       AAGC_net, A3GC_net, AGC_net, G_GRU_net,
                       PoseNet, PoseNet3, PoseNet_AGC, PoseNet_GGRU, pose_loss)
from .pipeline import TPPipeline, prepare_input, concat_stage_input, reduced_global_to_full_local, INPUT_JOINTS
from .sharding import shard_range, ShardedRunner, FlatGradAllReducer, train_step
from .train_loop import stage_inputs, checkpoint_name, latest_checkpoints, validate, fit_stage, teacher_forced_sample

__all__ = [n for n in dir() if not n.startswith("_")]
