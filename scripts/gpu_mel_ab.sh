#!/bin/bash
# mel kernel A/B: parity tests with the new build, then timing of every e2e_tts_b200/lib/ab_mel*.so against the new build
# (cfg5 shape, a 64-clip batch, one clip), interleaved twice on the same box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mel.py tests/test_gpu_full_size.py -q -m gpu -k "mel or Mel or stft or crop" > gpurun_out/mel_ab_pytest.log 2>&1; tail -3 gpurun_out/mel_ab_pytest.log
: > gpurun_out/mel_ab.log
for rep in 1 2; do
  for shape in "1024 220500 10" "64 220500 20" "1 110250 50"; do
    for lib in e2e_tts_b200/lib/ab_mel*.so; do
      echo "$(basename $lib .so) $shape: $(E2E_TTS_B200_LIB=$PWD/$lib python scripts/time_mel.py $shape)" | tee -a gpurun_out/mel_ab.log
    done
    echo "new $shape: $(python scripts/time_mel.py $shape)" | tee -a gpurun_out/mel_ab.log
  done
done
grep -E "mel" gpurun_out/parity_margins.jsonl | tail -12
