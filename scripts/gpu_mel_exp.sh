#!/bin/bash
# mel kernel: where the time goes (timing-only variants built with -DMEL_EXP=mask: 1 no filterbank pass, 2 no recombination /
# magnitudes, 4 no FFT passes 2-3, 8 no frame loop)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/mel_exp.log
echo "base: $(python scripts/time_mel.py 1024 220500 10)" | tee -a gpurun_out/mel_exp.log
for e in 1 2 3 4 7 8; do echo "MEL_EXP=$e: $(E2E_TTS_B200_LIB=$PWD/e2e_tts_b200/lib/mel_exp$e.so python scripts/time_mel.py 1024 220500 10)" | tee -a gpurun_out/mel_exp.log; done
