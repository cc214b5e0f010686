#!/bin/sh
# Lists the reference .c files that the reference's own per-directory CMakeLists.txt name
# (so stray drivers / legacy files its build ignores are ignored here too), minus the
# multiprecision ("mup") wrappers, which are off in the default configuration.
# usage: ref_sources.sh <reference src dir> <dir> [<dir> ...]
SRC="$1"; shift
for d in "$@"; do
  [ -f "$SRC/$d/CMakeLists.txt" ] || continue
  for f in "$SRC/$d"/*.c; do
    b=$(basename "$f")
    case "$b" in mup_*|*_mp.c|*_mp_device.c|umpire.c) continue;; esac
    if grep -qw "$b" "$SRC/$d/CMakeLists.txt"; then echo "$f"; fi
  done
done
