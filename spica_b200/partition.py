"""How work is split over GPUs. One process (or host thread) per GPU, scene
replicated, no data-path collective except the single film all-reduce per frame.

  samples : GPU g of G renders sample indices g, g+G, g+2G, ... (core/integrator.cc:64 is a loop of
            independent full-image passes; the sampler is counter-based, so the union over GPUs
            is exactly the single-GPU sample set)
  rays    : the ray-cast benchmark gives every GPU its own contiguous index range of the same
            counter-based ray generator (weak scaling)
"""


def sample_partition(spp, rank, world):
    """(first, count, stride) for spb_render_samples on `rank`."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    count = max(0, (spp - rank + world - 1) // world)
    return rank, count, world


def sample_indices(spp, rank, world):
    first, count, stride = sample_partition(spp, rank, world)
    return [first + i * stride for i in range(count)]


def ray_shard(n_per_gpu, rank):
    """(start, count) of the counter-based ray generator for `rank` (weak scaling)."""
    return rank * n_per_gpu, n_per_gpu
