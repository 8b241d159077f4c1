#!/usr/bin/env python
"""Tuning experiments only: run bench.py's main() against an experiment build of the library.

  python profiles/tools/bench_variant.py <path/to/libgcore_variant.so> [bench.py arguments]

The product package always loads ada-ray-tracer_b200/libgcore_b200.so; this wrapper points the already imported package at
another .so before its first lib() call.  Variant builds come from `make -C ada-ray-tracer_b200/csrc OUT=... OBJDIR=... VARIANT=...`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

lib = os.path.abspath(sys.argv[1])
sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[2:]
pkg = ge.load_package()
pkg.LIB_PATH = lib
import bench  # noqa: E402

bench.main()
