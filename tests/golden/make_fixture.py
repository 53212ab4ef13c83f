"""Extracts the only numeric fixture the reference ships: the brax HTML-viewer payload
embedded in /root/reference/notebooks/ant_tag.ipynb (cell 3, output 0): the complete
Ant(+Tag) brax Config as JSON plus a 21-frame rollout (pos[21,12,3], rot[21,12,4]).

Run in the build container (needs /root/reference); the outputs are committed:
  tests/golden/ant_tag_config.json   -- the reference's own Config (data, not code)
  tests/golden/ant_tag_rollout.npz   -- pos, rot (float64 as printed by the notebook)

    python tests/golden/make_fixture.py
"""
import json
import os
import re

import numpy as np

SRC = '/root/reference/notebooks/ant_tag.ipynb'
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    nb = json.load(open(SRC))
    html = ''.join(nb['cells'][3]['outputs'][0]['data']['text/html'])
    system = json.loads(re.search(r'var system = (\{.*?\});\s*\n', html, re.S).group(1))
    with open(os.path.join(HERE, 'ant_tag_config.json'), 'w') as f:
        json.dump(system['config'], f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, 'ant_tag_rollout.npz'),
                        pos=np.asarray(system['pos'], np.float64),
                        rot=np.asarray(system['rot'], np.float64))
    print('bodies', [b['name'] for b in system['config']['bodies']])


if __name__ == '__main__':
    main()
