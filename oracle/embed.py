"""TEST INFRASTRUCTURE: turn text files into C byte arrays (build step of oracle/_ref/regref)."""
import sys


def main():
    out = sys.argv[1]
    with open(out, "w") as f:
        for spec in sys.argv[2:]:
            name, path = spec.split("=", 1)
            data = open(path, "rb").read()
            f.write("static const unsigned char %s[] = {\n" % name)
            for i in range(0, len(data), 16):
                f.write("  " + ",".join(str(b) for b in data[i:i + 16]) + ",\n")
            f.write("  0\n};\nstatic const size_t %s_len = %d;\n" % (name, len(data)))


if __name__ == "__main__":
    main()
