#!/bin/bash
# Builds the library of another git branch (default: r2-candidates) into
# pfst_b200/csrc/libpfst_sm100_cand.so WITHOUT touching the working tree, so that one gpurun call
# can A/B it against the main build through PFST_LIB (pfst_b200/_lib.py). Run in the build container
# (needs .git); the .so is git-ignored and travels with the gpurun snapshot.
set -e
branch=${1:-r2-candidates}
root=$(cd "$(dirname "$0")/.." && pwd)
wt=$(mktemp -d /tmp/pfst_cand.XXXXXX)
git -C "$root" worktree add --detach "$wt" "$branch" > /dev/null
(cd "$wt" && python -m pfst_b200.build > /dev/null)
cp "$wt/pfst_b200/csrc/libpfst_sm100.so" "$root/pfst_b200/csrc/libpfst_sm100_cand.so"
git -C "$root" worktree remove --force "$wt"
echo "built $root/pfst_b200/csrc/libpfst_sm100_cand.so from $branch"
