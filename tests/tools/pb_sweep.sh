for v in pw3 pw4 pw6; do cp voice_synth_b200/lib/lib_$v.so voice_synth_b200/lib/libvoicesynth_cuda.so; echo $v; python tests/gpu_quick.py 2>&1 | grep -E "auto (synth|filter)" | cut -c1-90; done
