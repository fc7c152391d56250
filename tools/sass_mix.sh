#!/bin/bash
# usage: sass_mix.sh <cubin|so> <function-substring>   -> instruction histogram of matching kernels
cuobjdump -sass "$1" | awk -v pat="$2" '/Function :/{on=index($0,pat)>0; if(on) print} on && /^\s+\/\*[0-9a-f]{4}\*\//{print}' > /tmp/_sass_sel.txt
grep -c "^\s*/\*" /tmp/_sass_sel.txt
grep -oE "^\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P[0-9T]+ )?[A-Z0-9_.]+" /tmp/_sass_sel.txt | awk '{print $NF}' | sed -E 's/\.(reuse|FTZ|RN|SAT)//g' | sort | uniq -c | sort -rn | head -${3:-14}
