"""Differential fuzz (build container only: needs oracle/_ref/dwgsim_ref): random -m/-b/-v inputs through the compiled reference and the host shell with -M 2;
.mutations.txt/.vcf must be byte-identical, or both must fail (the reference's abort in mut_debug counts as a failure).
    python tools/fuzz_replay_vs_reference.py SEED N"""
import os, sys, random, subprocess, hashlib
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests','golden'))
import make_golden as mg
from oracle import pyoracle as po
ref=po.ref_binary(); cli=os.path.join(ROOT,'dwgsim_b200','bin','dwgsim')
WD='/tmp/dwgsim_fuzz'; os.makedirs(WD,exist_ok=True)
fa=WD+'/synth.fa'; mg.synth_fasta(fa)
contigs=[('chrA',30000),('chrB',12000),('tiny',400),('hp',8000)]
seqs={}
name=None
for line in open(fa):
    if line[0]=='>': name=line[1:].split()[0]; seqs[name]=[]
    else: seqs[name].append(line.strip())
seqs={k:''.join(v) for k,v in seqs.items()}
def md5(p): return hashlib.md5(open(p,'rb').read()).hexdigest() if os.path.exists(p) else None
def run(binary, args, prefix):
    r=subprocess.run([binary]+args+[fa,prefix],capture_output=True)
    return r.returncode, md5(prefix+'.mutations.txt'), md5(prefix+'.mutations.vcf'), r.stderr.decode(errors='ignore')[-200:]
rnd=random.Random(int(sys.argv[1]) if len(sys.argv)>1 else 1)
bad=0
for it in range(int(sys.argv[2]) if len(sys.argv)>2 else 100):
    kind=rnd.choice(['bed','txt','vcf'])
    path=WD+'/in.'+kind
    lines=[]
    for cname,clen in contigs:
        if rnd.random()<0.2: continue
        pos=rnd.randint(1,50)
        for _ in range(rnd.randint(0,12)):
            pos+=rnd.randint(1,clen//8)
            if pos>=clen-40: break
            if kind=='bed':
                t=rnd.choice(['snp','ins','del','SUB','I','d','insertion'])
                L=rnd.randint(1,rnd.choice([1,3,8,26]))
                end=pos+L
                bases=rnd.choice(['*',''.join(rnd.choice('ACGTacgtN') for _ in range(L))])
                lines.append('%s\t%d\t%d\t%s\t%s'%(cname,pos,end,bases,t))
                if rnd.random()<0.15: lines.append('%s\t%d\t%d\t%s\t%s'%(cname,pos,end,bases,t))  # overlap -> ignored
                pos=end
            elif kind=='txt':
                t=rnd.choice('SID')
                hap=rnd.choice([1,2,3])
                refb=seqs[cname][pos-1].upper()
                if refb not in 'ACGT': continue
                if t=='S':
                    alt=rnd.choice([b for b in 'ACGT' if b!=refb])
                    if hap<3:
                        codes="XACMGRSVTWYHKDBN"; alt=codes[(1<<'ACGT'.index(refb))|(1<<'ACGT'.index(alt))]
                    lines.append('%s\t%d\t%s\t%s\t%d'%(cname,pos,refb,alt,hap))
                elif t=='I':
                    ins=''.join(rnd.choice('ACGTN') for _ in range(rnd.randint(1,rnd.choice([2,10,40]))))
                    lines.append('%s\t%d\t-\t%s\t%d'%(cname,pos,ins,hap))
                else:
                    for k in range(rnd.randint(1,4)):
                        lines.append('%s\t%d\t%s\t-\t%d'%(cname,pos+k,seqs[cname][pos+k-1].upper(),hap))
                    pos+=4
            else:
                t=rnd.choice('SID'); tag=rnd.choice(['pl=1','pl=2','pl=3','AF=0.5;pl=1;mt=X','note'])
                refb=seqs[cname][pos-1:pos+5].upper()
                if any(c not in 'ACGT' for c in refb): continue
                if t=='S':
                    n=rnd.randint(1,3); alt=''.join(rnd.choice('ACGT') for _ in range(n))
                    lines.append('%s\t%d\t.\t%s\t%s\t.\t.\t%s'%(cname,pos,refb[:n],alt,tag))
                elif t=='I':
                    alt=refb[0]+''.join(rnd.choice('ACGT') for _ in range(rnd.randint(1,30)))
                    lines.append('%s\t%d\t.\t%s\t%s\t.\t.\t%s'%(cname,pos,refb[0],alt,tag))
                else:
                    n=rnd.randint(2,5)
                    lines.append('%s\t%d\t.\t%s\t%s\t.\t.\t%s'%(cname,pos,refb[:n],refb[0],tag))
                pos+=6
    if kind=='vcf': lines=['##fileformat=VCFv4.1','#CHROM\tPOS']+lines
    open(path,'w').write('\n'.join(lines)+'\n')
    flag={'bed':'-b','txt':'-m','vcf':'-v'}[kind]
    args=['-z',str(rnd.randint(1,99)),'-M','2',flag,path]+rnd.choice([[],['-H'],['-I','2','-X','0.5']])
    a=run(ref,args,WD+'/ref'); b=run(cli,args,WD+'/cli')
    for f in (WD+'/ref',WD+'/cli'):
        pass
    if a[0]==-6 and b[0]==1: a=b=(0,)
    if a[:3]!=b[:3]:
        bad+=1; print('MISMATCH',it,kind,args,a,b); 
        os.system('cp %s %s/bad_%d.%s'%(path,WD,it,kind))
        if bad>5: break
    for s in ('ref','cli'):
        for e in ('.mutations.txt','.mutations.vcf'):
            try: os.remove(WD+'/'+s+e)
            except FileNotFoundError: pass
print('done, mismatches',bad)
