#!/bin/sh
# usage: gen_stubs_all.sh obj...  -> C file on stdout with an abort() stub for every symbol the given objects reference but
# neither they nor the system libraries define (e.g. sonLib's stPhylogeny.c, which needs quicktree + spimap and is never
# reached from the BAR path). Test infrastructure.
TMP=$(mktemp -d)
echo 'int main(void){return 0;}' > $TMP/m.c
/usr/bin/gcc -fopenmp -o $TMP/a.out $TMP/m.c "$@" -lm -lz -lpthread 2> $TMP/err
echo '#include <stdio.h>'
echo '#include <stdlib.h>'
grep -o "undefined reference to \`[A-Za-z0-9_]*'" $TMP/err | sed "s/.*\`//; s/'//" | sort -u | while read s; do
  printf 'void %s(void) { fputs("oracle stub: %s called\\n", stderr); abort(); }\n' "$s" "$s"
done
rm -rf $TMP
