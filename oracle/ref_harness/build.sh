#!/usr/bin/env bash
# ORACLE / TEST INFRASTRUCTURE.  Builds oracle/_ref/rtiaw_ref_harness from the reference's own sources where they lie
# (/root/reference: path dependency, nothing is copied) with cargo --release.  This image has no Rust toolchain, so this script
# exits 3 here; tests/test_ref_harness.py skips on that.  With a toolchain (and the pinned crates of the reference's Cargo.lock in the
# local registry — there is no network) it produces the binary that PINS the oracle: the test then compares the C++ oracle with the
# real Rust Scene::hit, bit for bit.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../_ref"
if ! command -v cargo >/dev/null 2>&1; then
    echo "cargo not found: the reference (Rust) cannot be built in this image" >&2
    exit 3
fi
if [ ! -f /root/reference/Cargo.toml ]; then
    echo "/root/reference is not present (GPU box): nothing to build" >&2
    exit 3
fi
mkdir -p "$OUT"
cp /root/reference/Cargo.lock "$HERE/Cargo.lock" 2>/dev/null || true   # same pinned crate versions as the reference
CARGO_TARGET_DIR="$OUT/target" cargo build --release --offline --manifest-path "$HERE/Cargo.toml"
cp "$OUT/target/release/rtiaw_ref_harness" "$OUT/rtiaw_ref_harness"
echo "built $OUT/rtiaw_ref_harness"
