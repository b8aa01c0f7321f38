#!/bin/bash
# Round-2 hygiene pass (SURVEY 5 / VERDICT r1 item 10): compute-sanitizer memcheck + racecheck over small-shape GPU tests that reach every
# kernel family (SIMT ops, tcgen05 GEMM / attention forward+backward, decode launch chain, persistent and cluster decode engines, sampler).
# Output: gpurun_out/r02_sanitizer_*.log (summaries copied to profiles/).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TXL_TEST_FAST=1
SEL_SMALL='test_gemm_simt_shapes or test_embed_posemb or test_logsoftmax_nll or test_add_ln or test_adamw or test_io_formats or test_ntp_acc or (test_relattn_fwd_bwd and not 2048 and not 512 and not 320) or test_decode_tail_equals or test_dec_linear'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 0 python -m pytest tests/test_ops_gpu.py tests/test_tc_gemm_gpu.py -m gpu -q -x -k "$SEL_SMALL" > gpurun_out/r02_sanitizer_memcheck_ops.log 2>&1
tail -5 gpurun_out/r02_sanitizer_memcheck_ops.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 0 python -m pytest tests/test_parity_r2_gpu.py tests/test_model_gpu.py -m gpu -q -x -k "decode_step_bf16_vs_oracle or adaptive_softmax_losses or device_label_fixup or test_forward_loss_logits_mems or test_backward_grads or generate_sampling" > gpurun_out/r02_sanitizer_memcheck_model.log 2>&1
tail -5 gpurun_out/r02_sanitizer_memcheck_model.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 0 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "(test_relattn_fwd_bwd and 128-128-128-1024-1) or test_decode_tail_equals or test_add_ln" > gpurun_out/r02_sanitizer_racecheck_ops.log 2>&1
tail -5 gpurun_out/r02_sanitizer_racecheck_ops.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 0 python -m pytest tests/test_parity_r2_gpu.py -m gpu -q -x -k "decode_step_bf16_vs_oracle" > gpurun_out/r02_sanitizer_racecheck_decode.log 2>&1
tail -5 gpurun_out/r02_sanitizer_racecheck_decode.log
