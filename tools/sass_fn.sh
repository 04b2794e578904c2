#!/bin/bash
# usage: tools/sass_fn.sh <object> <substring of the mangled kernel name>  -> SASS of the first matching function
fn=$(cuobjdump -sass "$1" | grep "Function :" | grep "$2" | head -1 | awk '{print $3}')
cuobjdump -sass -fun "$fn" "$1"
