bash scripts/ab.sh
cp ab/K.so hyperreel_b200/libhyperreel_b200.so
bash scripts/gpu_check.sh
