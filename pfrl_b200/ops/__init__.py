"""torch.autograd wrappers around the fused CUDA kernels of libb2rl.so."""
