# static SASS instruction count per kernel of the built library:  bash profiles/sass_size.sh [regex]
cuobjdump -sass ${LIB:-demcoreg_b200/csrc/libdemcoreg_b200.so} | awk -v re="${1:-.}" '
/Function :/ { name=$3 }
/^ +\/\*[0-9a-f]{4}\*\/ / { n[name]++ }
END { for (k in n) if (k ~ re) printf "%6d instr %7.1f KB  %s\n", n[k], n[k]*16/1024, k }' | sort -n
