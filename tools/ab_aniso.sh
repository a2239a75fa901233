for d in build_variants/*/; do n=$(basename $d); echo $n; FLUIDMARCH_AB=1 FLUIDMARCH_LIB=$PWD/$d/libfluidmarch.so python tools/prof_step_aniso.py C2 4 | cut -c1-330; done
