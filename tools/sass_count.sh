#!/bin/bash
# instruction count of one kernel's SASS:  tools/sass_count.sh lib.so MANGLED_SUBSTRING
cuobjdump -sass "$1" 2>/dev/null | awk -v pat="$2" '/Function : /{on=index($0,pat)>0; if(on)name=$3} on && /^ +\/\*[0-9a-f]+\*\/ /{c++} END{print c, name}'
