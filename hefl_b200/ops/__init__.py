"""Python wrappers around the hand-written sm_100a kernels (torch.ops.hefl.*)."""
