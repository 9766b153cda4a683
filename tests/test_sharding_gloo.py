"""N>1 host logic on CPU: world_size-2 gloo processes (image sharding, single weight broadcast, gather)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from paint_with_words_sd_b200 import sharding
from paint_with_words_sd_b200.unet import UNetConfig, build_unet


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.set_num_threads(1)
    assert sharding.init_distributed("gloo")
    cfg = UNetConfig(block_out_channels=(32, 64, 64, 64), cross_attention_dim=16, attention_heads=2, norm_num_groups=8)
    unet = build_unet(cfg, seed=100 + rank)                 # deliberately different per rank before the broadcast
    nbytes = sharding.broadcast_module_weights(unet, src=0)
    ref = build_unet(cfg, seed=100)
    same = all(torch.equal(a, b) for a, b in zip(unet.parameters(), ref.parameters()))
    mine = sharding.shard_images(5, rank, world)
    local = {i: torch.full((1, 4, 2, 2), float(i)) for i in mine}
    allv = sharding.gather_latents(local, 5)
    ok_gather = all(float(allv[i].mean()) == float(i) for i in range(5))
    ret[rank] = (nbytes, same, mine, ok_gather)
    dist.destroy_process_group()


def test_world2_broadcast_shard_gather():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        r0, r1 = ret[0], ret[1]
    assert r0[0] == r1[0] > 0
    assert r0[1] and r1[1]                                   # replicas identical to rank 0's weights
    assert r0[2] == [0, 2, 4] and r1[2] == [1, 3]            # image i -> rank i mod G
    assert r0[3] and r1[3]


def test_single_process_is_noop():
    unet = build_unet(UNetConfig(block_out_channels=(32, 64, 64, 64), cross_attention_dim=16, attention_heads=2,
                                 norm_num_groups=8), seed=0)
    assert sharding.broadcast_module_weights(unet) == 0
    assert sharding.shard_images(3, 0, 1) == [0, 1, 2]
