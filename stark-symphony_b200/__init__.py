"""stark-symphony_b200 — B200-native batched verifier for the stark-symphony programs.

Host-side mirror of the reference's interface for its verification hot path:

  reference                                              here
  ------------------------------------------------------ ----------------------------------------------
  `simfony run prog.simf --witness proof.wit`            Verifier.run_stwo_wit / run_stark101_wit, bin/verify-batch
   (simfony-cli/src/main.rs:163-209)
  stwo-verifier/scripts/generate_wit.py                  witness.stwo_wit_from_proof_json, witness.pack_stwo_proof_json
  stark101/scripts/generate_wit.py                       witness.stark101_wit_from_proof_json, witness.pack_stark101_proof_json
  each `.simf` fn / jet                                  Verifier.m31_mul, qm31_mul, circle_fold, sha256_pair, merkle_verify_32, ...

All compute goes through libssym.so (hand-written sm_100a CUDA behind the C-ABI of include/ssym.h).
Nothing in this package imports oracle/: that directory is the tests' checker only.
"""
from ._lib import (ERR_CUDA, ERR_INTERNAL, ERR_NOMEM, ERR_PARSE, ERR_USAGE, MEM_DEVICE, MEM_HOST, MODE_PROVER_CONSISTENT, MODE_QUERY_DEDUP, MODE_REF_LITERAL, S101Trace, SsymError, StwoConfig, StwoLayout, StwoTrace,
                   load)
from .verifier import Verifier, stwo_config, stwo_layout
from . import witness

__all__ = ["Verifier", "stwo_config", "stwo_layout", "witness", "StwoConfig", "StwoLayout", "StwoTrace", "S101Trace", "SsymError", "load",
           "MEM_DEVICE", "MEM_HOST", "MODE_REF_LITERAL", "MODE_PROVER_CONSISTENT", "MODE_QUERY_DEDUP", "ERR_USAGE", "ERR_CUDA", "ERR_PARSE", "ERR_NOMEM", "ERR_INTERNAL"]
