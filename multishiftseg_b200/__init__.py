"""multishiftseg_b200 -- B200-native dense anomaly scoring + exact OOD evaluation.

Drop-in for the scoring/evaluation hot path of gaozhitong/MultiShiftSeg:

    from multishiftseg_b200.metric import eval_ood_measure           # lib/utils/metric.py:170
    from multishiftseg_b200.deeplab import energy_func, Upsample      # deepv3.py:251, mynn.py:28
    from multishiftseg_b200.m2f import semantic_inference, get_anomaly_score

Everything runs in hand-written sm_100a CUDA kernels behind ``libmss_b200.so`` (C ABI in
``include/mss_b200.h``); importing this package does not load the library, the first call does, and it
raises if the library has not been built -- there is no CPU or PyTorch fallback.
"""
__version__ = "0.1.0"
