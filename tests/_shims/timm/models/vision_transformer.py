from torch.nn.init import trunc_normal_  # noqa: F401  (same signature: mean, std, a, b)
