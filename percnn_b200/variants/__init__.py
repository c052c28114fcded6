"""One module per reference script family; each exports `RCNNCell`, `RCNN` (and `upscaler`) with the
constructor signatures of that script, backed by the fused CUDA cell.

    lambda_omega_fwd  ForwardSimulationOfPDEs/2d_lambda_omega/percnn_LO_eqn.py        (fp64, k=1, hc=4)
    gs2d              DataDrivenModeling/2d_gs_rd/train_2drd.py                       (fp32, k=1, hc=8)
    gs3d              DataDrivenModeling/3d_gs_rd/train_3drd.py                       (fp32, k=1, hc=2)
    burgers_stage1    DataDrivenDiscoveryOfPDEs/2D_Burgers_eqn/Stage-1/rcnn_Burgers_* (fp32, k=5, hc=16)
    lo_stage1         DataDrivenDiscoveryOfPDEs/2D_Lambda_Omega_eqn/stage-1/rcnn_LO_* (fp32, k=5, hc=16)
    burgers_stage3    DataDrivenDiscoveryOfPDEs/2D_Burgers_eqn/Stage-3/fine_tuning_*  (fp64, physics RHS)
    lo_stage3         DataDrivenDiscoveryOfPDEs/2D_Lambda_Omega_eqn/stage-3/fine_tuning_LO_* (fp64)
"""
