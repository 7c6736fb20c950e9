#!/bin/bash
# Host-side memcheck / racecheck of the kernels that run on the CUDA emulator (tests/emu): builds tests/emu/sanitize_kernels.cpp
# with AddressSanitizer + UBSan and with ThreadSanitizer and runs both; a control kernel with a missing barrier shows that
# ThreadSanitizer sees races between emulated CUDA threads.  No GPU needed.  Usage: bash scripts/emu_sanitize.sh [logdir]
set -u
cd "$(dirname "$0")/../tests/emu" || exit 1
mkdir -p _build
out=${1:-_build}
FLAGS="-std=c++20 -O1 -g -pthread -ffp-contract=off -Wno-tsan"
g++ $FLAGS -fsanitize=address,undefined -fno-sanitize-recover=undefined -o _build/sanitize_asan sanitize_kernels.cpp || exit 1
g++ $FLAGS -fsanitize=thread -o _build/sanitize_tsan sanitize_kernels.cpp || exit 1
g++ $FLAGS -fsanitize=thread -o _build/race_bad race_selftest.cpp || exit 1
g++ $FLAGS -fsanitize=thread -DWITH_BARRIER -o _build/race_ok race_selftest.cpp || exit 1
ASAN_OPTIONS=detect_leaks=0 ./_build/sanitize_asan > "$out/emu_asan.log" 2>&1; a=$?
ASAN_OPTIONS=detect_leaks=0 ./_build/sanitize_asan denoise >> "$out/emu_asan.log" 2>&1; a=$((a + $?))
ASAN_OPTIONS=detect_leaks=0 ./_build/sanitize_asan general >> "$out/emu_asan.log" 2>&1; a=$((a + $?))
ASAN_OPTIONS=detect_leaks=0 ./_build/sanitize_asan resize >> "$out/emu_asan.log" 2>&1; a=$((a + $?))
ASAN_OPTIONS=detect_leaks=0 ./_build/sanitize_asan dense >> "$out/emu_asan.log" 2>&1; a=$((a + $?))
TSAN_OPTIONS="halt_on_error=0" ./_build/sanitize_tsan > "$out/emu_tsan.log" 2>&1; t=$?
TSAN_OPTIONS="halt_on_error=0" ./_build/sanitize_tsan denoise >> "$out/emu_tsan.log" 2>&1; t=$((t + $?))
TSAN_OPTIONS="halt_on_error=0" ./_build/sanitize_tsan general >> "$out/emu_tsan.log" 2>&1; t=$((t + $?))
TSAN_OPTIONS="halt_on_error=0" ./_build/sanitize_tsan resize >> "$out/emu_tsan.log" 2>&1; t=$((t + $?))
TSAN_OPTIONS="halt_on_error=0" ./_build/sanitize_tsan dense >> "$out/emu_tsan.log" 2>&1; t=$((t + $?))
./_build/race_bad > "$out/emu_tsan_control_bad.log" 2>&1
./_build/race_ok > "$out/emu_tsan_control_ok.log" 2>&1
echo "asan+ubsan: exit $a, $(grep -cE 'ERROR: AddressSanitizer|runtime error' "$out/emu_asan.log") reports"
echo "tsan: exit $t, $(grep -c 'WARNING: ThreadSanitizer' "$out/emu_tsan.log") reports"
echo "tsan control without the barrier: $(grep -c 'WARNING: ThreadSanitizer: data race' "$out/emu_tsan_control_bad.log") race report(s) (expected >= 1)"
echo "tsan control with the barrier: $(grep -c 'WARNING: ThreadSanitizer' "$out/emu_tsan_control_ok.log") report(s) (expected 0)"
tail -n 1 "$out/emu_asan.log"; tail -n 1 "$out/emu_tsan.log"
[ $a -eq 0 ] && [ $t -eq 0 ]
