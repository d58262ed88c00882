"""Differential fuzz (build container only: needs oracle/_ref/dwgsim_ref): random option sets (-r -R -X -I -H -1 -2 -d -s -c -x ...) through the compiled reference and the
host shell with -M 2 / -C 0; .mutations.txt/.vcf must be byte-identical.
    python tools/fuzz_mutations_vs_reference.py SEED N"""
import os, sys, random, subprocess, hashlib
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests','golden'))
import make_golden as mg
from oracle import pyoracle as po
ref=po.ref_binary(); cli=os.path.join(ROOT,'dwgsim_b200','bin','dwgsim')
WD='/tmp/dwgsim_fuzz'; os.makedirs(WD,exist_ok=True)
fa=WD+'/synth.fa'; mg.synth_fasta(fa)
if len(sys.argv)>3 and sys.argv[3]=='random-fasta':
    sys.path.insert(0,os.path.join(ROOT,'tools'))
    import importlib.util
    spec=importlib.util.spec_from_file_location('fo',os.path.join(ROOT,'tools','fuzz_oracle_vs_reference.py'))
    src=open(os.path.join(ROOT,'tools','fuzz_oracle_vs_reference.py')).read()
    ns={}; exec(src[src.index('def random_fasta'):src.index('fa = mg.synth_fasta')],ns)
    fa=ns['random_fasta'](WD+'/rand_m.fa',int(sys.argv[1]))
def md5(p): return hashlib.md5(open(p,'rb').read()).hexdigest() if os.path.exists(p) else None
def run(binary, args, prefix):
    for e in ('.mutations.txt','.mutations.vcf'):
        try: os.remove(prefix+e)
        except FileNotFoundError: pass
    r=subprocess.run([binary]+args+[fa,prefix],capture_output=True)
    return r.returncode, md5(prefix+'.mutations.txt'), md5(prefix+'.mutations.vcf'), r.stderr.decode(errors='ignore')[-160:]
rnd=random.Random(int(sys.argv[1])); bad=0
for it in range(int(sys.argv[2])):
    args=['-z',str(rnd.randint(0,10**6))]
    args+=rnd.choice([['-M','2'],['-C','0']])
    if rnd.random()<0.7: args+=['-r',rnd.choice(['0','0.0001','0.001','0.01','0.1','0.5'])]
    if rnd.random()<0.6: args+=['-R',rnd.choice(['0','0.1','0.5','0.9','1'])]
    if rnd.random()<0.5: args+=['-X',rnd.choice(['0','0.3','0.7','0.95','0.99'])]
    if rnd.random()<0.4: args+=['-I',rnd.choice(['1','2','5','30'])]
    if rnd.random()<0.3: args+=['-H']
    if rnd.random()<0.3: args+=['-1',rnd.choice(['50','100','250']),'-2',rnd.choice(['0','50','100'])]
    if rnd.random()<0.3: args+=['-d',rnd.choice(['200','500','3000']),'-s',rnd.choice(['10','50','500'])]
    if rnd.random()<0.2: args+=['-c',rnd.choice(['0','1'])]
    if rnd.random()<0.25:
        bed=WD+'/reg.bed'; lines=[]
        for cname,clen in (('chrA',30000),('chrB',12000),('hp',8000)):
            if rnd.random()<0.3: continue
            p=rnd.randint(1,2000)
            while p<clen-600 and rnd.random()<0.8:
                e=min(clen,p+rnd.randint(300,5000)); lines.append('%s\t%d\t%d'%(cname,p,e)); p=e+rnd.randint(-100,3000)
                if p<1: p=1
        open(bed,'w').write('\n'.join(lines)+'\n')
        if lines: args+=['-x',bed]
    a=run(ref,args,WD+'/ref'); b=run(cli,args,WD+'/cli')
    if a[:3]!=b[:3]:
        bad+=1; print('MISMATCH',it,args,a,b)
        if bad>4: break
print('done, mismatches',bad)
