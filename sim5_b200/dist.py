"""Row-split helpers for one-process-per-GPU runs (torch.distributed; NCCL on GPUs, gloo in CPU tests).

The photon path shards by rows: every ray is independent, per-image constants are ~0.7 KB.  Rank r of W
traces the row blocks b (of `block_rows` rows) with b % W == r  -- an INTERLEAVED static split, so the
expensive rows around the black-hole shadow are dealt round-robin to all ranks (per-rank cost equal to
within ~1 %), which is what the reference's "static + work stealing" intent needs without any
cross-process counter.  The only collective is the gather of the finished planes at the end (images) or the reduce(sum)
of the per-rank transfer-function histograms (HISTOGRAM mode, lattice_images).
"""

DEFAULT_BLOCK_ROWS = 32


def check_split(ny, world, block_rows=DEFAULT_BLOCK_ROWS):
    if world < 1 or block_rows < 1 or ny % (world * block_rows) != 0:
        raise ValueError("ny=%d must be a multiple of world*block_rows=%d*%d" % (ny, world, block_rows))
    return ny // world


def apply_split(p, rank, world, block_rows=DEFAULT_BLOCK_ROWS):
    """Set the split fields of a sim5_image_params for `rank` of `world`; returns the local row count."""
    rows = check_split(p.ny, world, block_rows)
    p.split_count, p.split_index, p.split_rows = world, rank, block_rows
    return rows


def local_rows(ny, rank, world, block_rows=DEFAULT_BLOCK_ROWS):
    """Image rows traced by `rank`, in local-row order (what a compact device plane holds)."""
    rows = check_split(ny, world, block_rows)
    out = []
    for lr in range(rows):
        out.append(((lr // block_rows) * world + rank) * block_rows + lr % block_rows)
    return out


def assemble(parts, world, block_rows=DEFAULT_BLOCK_ROWS):
    """Full image [ny, nx] from the per-rank compact planes [rows_local, nx] (numpy arrays or torch tensors,
    in rank order): stack -> [W, nb, block_rows, nx] -> swap the first two axes -> [ny, nx]."""
    first = parts[0]
    rows_local, nx = first.shape
    nb = rows_local // block_rows
    if hasattr(first, "permute"):      # torch
        import torch
        st = torch.stack(list(parts), 0).view(world, nb, block_rows, nx)
        return st.permute(1, 0, 2, 3).reshape(world * rows_local, nx)
    import numpy as np
    st = np.stack(list(parts), 0).reshape(world, nb, block_rows, nx)
    return st.transpose(1, 0, 2, 3).reshape(world * rows_local, nx)


def lattice_images(n_images, rank, world):
    """The (spin, inclination) lattice images of HISTOGRAM mode that `rank` traces: an INTERLEAVED deal, image i -> rank i mod W
    (what split_count / split_index mean in that mode).  Every rank zeroes the bins of the other ranks' images, so the per-rank
    histograms add up to the lattice: one reduce(sum) at the end."""
    return list(range(rank, n_images, world))


def lattice_range(n_images, rank, world):
    """Contiguous share of the (spin, inclination) lattice of the HISTOGRAM mode for `rank`."""
    per = (n_images + world - 1) // world
    lo = min(rank * per, n_images)
    hi = min(lo + per, n_images)
    return lo, hi
