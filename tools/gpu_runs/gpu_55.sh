# config 5 at full size on 8 GPUs with the current guiding-centre kernel (the r1 number predates the separable-field change)
bash tools/gpu_multi_gc2.sh 8 belt 12500000 10
