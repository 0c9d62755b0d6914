import torch


def create_meshgrid3d(depth, height, width, normalized_coordinates=True, device=torch.device("cpu"),
                      dtype=torch.float32):
    """kornia 0.5.0 semantics: returns (1, D, H, W, 3) with last dim = (x:width, y:height, z:depth)...
    kornia 0.5.0 stacks meshgrid([zs, xs, ys]) as (D, W, H, 3) then permutes(0,2,1) -> (1, D, H, W, 3)
    with the last dimension ordered (z-index over depth, x-index over width, y-index over height)."""
    xs = torch.linspace(0, width - 1, width, device=device, dtype=dtype)
    ys = torch.linspace(0, height - 1, height, device=device, dtype=dtype)
    zs = torch.linspace(0, depth - 1, depth, device=device, dtype=dtype)
    if normalized_coordinates:
        xs = (xs / (width - 1) - 0.5) * 2
        ys = (ys / (height - 1) - 0.5) * 2
        zs = (zs / (depth - 1) - 0.5) * 2
    base_grid = torch.stack(torch.meshgrid([zs, xs, ys], indexing="ij"), dim=-1)  # D x W x H x 3
    return base_grid.permute(0, 2, 1, 3).unsqueeze(0)  # 1 x D x H x W x 3
