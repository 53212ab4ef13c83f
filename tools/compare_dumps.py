import sys, torch
a, b = torch.load(sys.argv[1]), torch.load(sys.argv[2])
for kind in a:
    for k in a[kind]:
        same = torch.equal(a[kind][k], b[kind][k])
        d = (a[kind][k].double() - b[kind][k].double()).abs().max().item()
        print(kind, k, 'bit-identical' if same else f'DIFFERENT max |d| {d:.3e}, {int((a[kind][k] != b[kind][k]).sum())} entries')
