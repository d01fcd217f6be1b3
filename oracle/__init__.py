"""CPU oracle for the ContrastiveProsthetics hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithm of the reference's hot path
(/root/reference/code: load.py DB23, utils.py TaskWrapper, models.py Model /
EMGNet / GLOVENet, plus the README-only class-subset evaluator).  Each function
cites the reference file:line it follows.

Who may import it: tests/, __graft_entry__.smoke() (as the checker) and
bench.py's cpu_baseline / --impl reference leg (as the timed CPU baseline).
The product package `contrastiveprosthetics_b200` never imports it and has no
CPU fallback: without the CUDA library it raises.

Pinning ("how do we know the oracle is the reference?"):
  * tests/golden/make_golden.py imports the UNMODIFIED reference modules in
    the build container (CPU, two mechanical shims) and stores their outputs
    on seeded synthetic NinaPro-shaped inputs in tests/golden/*.npz;
    tests/test_oracle_golden.py replays every fixture through this oracle.
  * the reference's own result artefacts (data/y_pred.npy, y_true.npy,
    voting.npy, confusion_matrix.npy) are checked through the three
    known-answer relations of SURVEY.md section 4.
  * the class-subset evaluator (README.md:11,15) has NO runnable reference:
    its restatement is "parity unpinned" beyond (a) full-set subset ==
    Model.loss's vote/y_pred and (b) structural relations of data/*_grasp.xlsx.

Arithmetic: float work is torch-CPU fp32 (the reference's own arithmetic
library) with an fp64 switch for error-bar context; integer work (argmax,
vote, counts) is numpy, with a plain-C twin (oracle/vote_subset.c) used for
the full-size CPU baseline.
"""
