"""hiertcn_b200 -- B200-native (sm_100a) implementation of the HierTCN hot path.

Host-side mirror of the reference surface (numpy / torch plumbing only; all arithmetic is in libhtcn.so):

    args.make_args            hyper-parameters with the reference's flag names / defaults   (reference args.py)
    data_loader               dequeue() batch layout, queue loader, synthetic XING-shaped data (reference data_loader.py)
    weights                   state dict keyed by TF variable names, glorot init, npz io       
    model_hier.HierTCN        build / forward / loss / step / step_async, model_hier(...)      (reference model_hier.py, model.py)
    model_tcn.TCN             single-level TCN, model_tcn(...)                                  (reference model_tcn.py)
    loss                      calc_loss / calc_score / calc_metric_fast / top_k                 (reference loss.py)
    run_hier.evaluate_hier    evaluation loop with device-resident carried state                (reference run_hier_xing.py)
    train.HierTCNTrainer      training step: forward + backward + TF-flavoured Adam, run_hier loop (model.py:134-141)
    dist                      data-parallel scalar all-reduce, catalog-sharded scoring (NCCL)
    _cabi                     ctypes binding of include/htcn.h

Importing the package never touches CUDA; the library is loaded on first use and there is no CPU fallback.
"""
__version__ = "0.1.0"
