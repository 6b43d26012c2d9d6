#!/bin/sh
# usage: gen_stubs.sh main.o "other objects"  -> C file on stdout with an abort() stub for every symbol
# that main.o needs and neither the other objects nor the system libraries provide (the Cactus graph API
# used only by the Flower-level functions of poaBarAligner.c, which the MSA-level checkers never call).
MAIN="$1"; OTHERS="$2"
TMP=$(mktemp -d)
echo 'int main(void){return 0;}' > $TMP/m.c
/usr/bin/gcc -fopenmp -o $TMP/a.out $TMP/m.c -Wl,--whole-archive "$MAIN" -Wl,--no-whole-archive $OTHERS -lm -lz -lpthread 2> $TMP/err
echo '#include <stdio.h>'
echo '#include <stdlib.h>'
grep -o "undefined reference to \`[A-Za-z0-9_]*'" $TMP/err | sed "s/.*\`//; s/'//" | sort -u | while read s; do
  printf 'void %s(void) { fputs("oracle stub: %s called", stderr); abort(); }\n' "$s" "$s"
done
rm -rf $TMP
