"""The B200 backend: schedule (fused row-streaming stages), CUDA emission, host-class emission."""
