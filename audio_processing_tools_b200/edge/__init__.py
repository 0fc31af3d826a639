"""GPU twins of the reference's `edge` package: the spectral noise / rain detector engine, the band noise estimator
and the legacy RoE detector.  Every module here calls the CUDA library through `_lib`; none computes on the CPU."""
