#!/bin/bash
# usage: tools/sass_mix.sh <mangled-substring>   -> opcode histogram of one kernel in the built library
cuobjdump -sass vmambair_b200/lib/libvmambair_b200.so | awk '/Function : /{name=$3} {print name "\t" $0}' | grep "$1" | grep -oE "^\S+\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T] +)?[A-Z0-9_.]+" | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -${2:-16}
