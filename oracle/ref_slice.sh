#!/bin/sh
# ORACLE — test infrastructure only.  Cuts lines START..END of a reference source file into oracle/_ref/gen/ so that member
# functions living in translation units too entangled to compile whole (Tracking.cc, MapPoint.cc, KeyFrame.cc) are still
# compiled from the reference's own text.  Fails loudly when the first / last line of the cut is not the expected one.
# usage: ref_slice.sh SRC START END FIRST_REGEX LAST_REGEX OUT
set -e
src="$1"; a="$2"; b="$3"; first="$4"; last="$5"; out="$6"
sed -n "${a}p" "$src" | grep -Eq "$first" || { echo "ref_slice: $src:$a is not /$first/" >&2; exit 1; }
sed -n "${b}p" "$src" | grep -Eq "$last" || { echo "ref_slice: $src:$b is not /$last/" >&2; exit 1; }
mkdir -p "$(dirname "$out")"
{ echo "/* GENERATED at build time by oracle/ref_slice.sh: $src lines $a-$b, verbatim.  Not tracked. */"; echo "#line $a \"$src\""; sed -n "${a},${b}p" "$src"; } > "$out"
