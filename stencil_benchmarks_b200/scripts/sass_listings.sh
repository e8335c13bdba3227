#!/bin/bash
# cuobjdump -sass listings of the shipped f64 kernels into profiles/sass/ (runs without a GPU).
set -e
lib=stencil_benchmarks_b200/csrc/libsbench_b200.so
out=profiles/sass
mkdir -p $out
dump() {  # $1 = output name, $2 = grep pattern on the mangled name
  name=$(cuobjdump -sass $lib 2>/dev/null | grep "Function :" | sed 's/.*Function : //' | grep -E "$2" | head -1)
  [ -n "$name" ] || { echo "no kernel matches $2"; return; }
  cuobjdump -sass -fun "$name" $lib > $out/$1.sass 2>/dev/null
  echo "$1: $(c++filt "$name" | cut -c1-110)  ($(grep -cE '^\s+/\*[0-9a-f]{4}\*/' $out/$1.sass) instructions)"
}
dump hdiff_tma_kernel_f64 'hdiff_tma_kernelIdLi4ELi4ELb0'
dump hdiff_tma_kernel_peer_f64 'hdiff_tma_kernelIdLi4ELi4ELb1'
dump hdiff_jmarch_kernel_f64 'hdiff_jmarch_kernelIdLi2ELi64'
dump vadv_onchip_kernel_f64 'vadv_onchip_kernelIdLi4'
dump vadv_global_kernel_f64 'vadv_global_kernelId'
dump basic_kernel_f64 'basic_kernelIdLi4ELi3ELi2'
dump stream_kernel_f64 'stream_kernelIdLi3ELi16ELi4ELb1'
