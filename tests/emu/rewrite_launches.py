"""kernel<<<grid, block, smem, stream>>>(args)  ->  emu::launch(kernel, grid, block, smem, args)   (tests/emu only)."""
import sys


def balanced(s, i, open_ch, close_ch):
    depth = 0
    for j in range(i, len(s)):
        if s[j] == open_ch:
            depth += 1
        elif s[j] == close_ch:
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced")


def rewrite(src):
    out, pos = [], 0
    while True:
        i = src.find("<<<", pos)
        if i < 0:
            out.append(src[pos:])
            return "".join(out)
        # kernel expression: identifier, optionally followed by <template args>
        k0 = i
        if src[k0 - 1] == ">":
            depth, k0 = 0, i - 1
            while True:
                if src[k0] == ">":
                    depth += 1
                elif src[k0] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                k0 -= 1
        while k0 > 0 and (src[k0 - 1].isalnum() or src[k0 - 1] in "_:"):
            k0 -= 1
        j = src.find(">>>", i)
        cfg = [c.strip() for c in src[i + 3:j].split(",")]
        assert len(cfg) == 4, cfg
        p0 = src.index("(", j)
        p1 = balanced(src, p0, "(", ")")
        out.append(src[pos:k0])
        out.append(f"emu::launch({src[k0:i]}, {cfg[0]}, {cfg[1]}, (size_t)({cfg[2]}), {src[p0 + 1:p1]})")
        pos = p1 + 1


if __name__ == "__main__":
    sys.stdout.write(rewrite(open(sys.argv[1]).read()))
